// comm.cu — the two collectives of image-sharded multi-GPU inference, behind the C API (BASELINE north star:
// "weights are broadcast once over NVLink with NCCL, and detections are gathered only at the end, with no collective on the
// per-layer path").  One process per GPU; every process parses the cfg, rank `root` alone reads the .weights file, and
//   b200_comm_broadcast_weights   replicates the folded / repacked parameter arena with ONE ncclBroadcast,
//   b200_comm_set_gather          makes every later b200_detect_* call deliver all ranks' detection records to the root:
//                                 fixed-size slots (header + records) move with ncclSend / ncclRecv on the result stream, i.e.
//                                 beside the next batch's forward pass, never between layers.
// NCCL is resolved at run time (dlopen of libnccl.so.2, the copy the process already holds if any), so the library keeps
// loading on machines without it; the calls fail loudly when it is missing.  The reference has no multi-GPU inference at all
// (its `-gpus` option is training only, detector.c:20-60); this replaces nothing and adds the sharding the metric asks for.
#include "engine.h"
#include <dlfcn.h>
#include <cstring>
#include <nccl.h>

struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char *names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return nullptr;
#define B200_NCCL_SYM(field, sym) *(void **)(&api.field) = dlsym(api.lib, sym); if (!api.field) { api.lib = nullptr; return nullptr; }
    B200_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    B200_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    B200_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    B200_NCCL_SYM(Broadcast, "ncclBroadcast")
    B200_NCCL_SYM(AllReduce, "ncclAllReduce")
    B200_NCCL_SYM(Send, "ncclSend")
    B200_NCCL_SYM(Recv, "ncclRecv")
    B200_NCCL_SYM(GroupStart, "ncclGroupStart")
    B200_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    B200_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_NCCL_SYM
    return &api;
}

static bool comm_trace() { static int on = -1; if (on < 0) on = getenv("B200_COMM_TRACE") ? 1 : 0; return on == 1; }
#define B200_COMM_TRACE(...) do { if (comm_trace()) { fprintf(stderr, "[b200-comm] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)

static NcclApi *need_nccl(const char *what)
{
    NcclApi *a = nccl_api();
    if (!a) { fprintf(stderr, "b200-darknet: %s needs NCCL (libnccl.so.2 not found; set B200_NCCL_LIB)\n", what); abort(); }
    return a;
}

#define B200_NCCL_CHECK(api, call)                                                                                   \
    do {                                                                                                             \
        ncclResult_t r_ = (call);                                                                                    \
        if (r_ != ncclSuccess) {                                                                                     \
            fprintf(stderr, "b200-darknet: NCCL error %s at %s:%d\n", (api)->GetErrorString(r_), __FILE__, __LINE__); \
            abort();                                                                                                 \
        }                                                                                                            \
    } while (0)

struct B200Comm {
    ncclComm_t comm;
    int rank, world;
    int gather_root;           // -1: detections stay on their rank
    int slot_records;          // records per rank that travel to the root each batch
    int image_base;            // global number of this rank's image 0
    int *d_headers;            // root: [world][4] ints {records, image base, 0, 0}
    DetRecord *d_slots;        // root: [world][slot_records]
    int *h_headers;            // pinned
};

extern "C" int b200_comm_unique_id(unsigned char *id, int bytes)
{
    NcclApi *a = need_nccl("b200_comm_unique_id");
    if (bytes < (int)sizeof(ncclUniqueId)) return -1;
    ncclUniqueId u;
    B200_NCCL_CHECK(a, a->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return (int)sizeof u;
}

extern "C" int b200_comm_init(network *net, const unsigned char *id, int rank, int world)
{
    b200_engine *e = b200_engine_of(net);
    NcclApi *a = need_nccl("b200_comm_init");
    if (e->device < 0) { fprintf(stderr, "b200-darknet: b200_comm_init needs a CUDA device\n"); abort(); }
    B200_CHECK(cudaSetDevice(e->device));
    if (e->comm) b200_comm_release(e);
    B200Comm *c = new B200Comm();
    memset(c, 0, sizeof *c);
    c->rank = rank; c->world = world; c->gather_root = -1;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    B200_COMM_TRACE("rank %d/%d: ncclCommInitRank on device %d", rank, world, e->device);
    B200_NCCL_CHECK(a, a->CommInitRank(&c->comm, world, u, rank));
    B200_COMM_TRACE("rank %d: communicator ready", rank);
    e->comm = c;
    return 0;
}

extern "C" int b200_comm_rank(network *net) { b200_engine *e = b200_engine_of(net); return e->comm ? e->comm->rank : 0; }
extern "C" int b200_comm_world(network *net) { b200_engine *e = b200_engine_of(net); return e->comm ? e->comm->world : 1; }

// the parameter arena of `root` (filled by load_weights there) replaces this rank's: one collective, then the engine is
// ready on every rank.  Every rank must have parsed the same cfg at the same precision (same arena layout).
extern "C" int b200_comm_broadcast_weights(network *net, int root)
{
    b200_engine *e = b200_engine_of(net);
    if (!e->comm) { fprintf(stderr, "b200-darknet: b200_comm_broadcast_weights before b200_comm_init\n"); abort(); }
    NcclApi *a = need_nccl("b200_comm_broadcast_weights");
    B200_CHECK(cudaSetDevice(e->device));
    // every rank must hold the same plan (same cfg, same precision): a size mismatch would leave the broadcast hanging, so it
    // is checked first with a tiny all-reduce: max(bytes) and max(-bytes) agree only when all sizes are equal
    {
        long long h[2] = {(long long)e->arena_bytes, -(long long)e->arena_bytes}, *d = nullptr;
        B200_CHECK(cudaMalloc(&d, sizeof h));
        B200_CHECK(cudaMemcpyAsync(d, h, sizeof h, cudaMemcpyHostToDevice, e->stream));
        B200_NCCL_CHECK(a, a->AllReduce(d, d, 2, ncclInt64, ncclMax, e->comm->comm, e->stream));
        B200_CHECK(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, e->stream));
        B200_CHECK(cudaStreamSynchronize(e->stream));
        cudaFree(d);
        if (h[0] != -h[1]) {
            fprintf(stderr, "b200-darknet: b200_comm_broadcast_weights: rank %d holds a %zu-byte parameter arena, another rank one of %lld bytes — "
                            "every rank must parse the same cfg at the same precision\n", e->comm->rank, e->arena_bytes, h[0] == (long long)e->arena_bytes ? -h[1] : h[0]);
            abort();
        }
    }
    B200_COMM_TRACE("rank %d: broadcast of %zu parameter bytes from rank %d", e->comm->rank, e->arena_bytes, root);
    B200_NCCL_CHECK(a, a->Broadcast(e->arena, e->arena, e->arena_bytes, ncclUint8, root, e->comm->comm, e->stream));
    B200_COMM_TRACE("rank %d: broadcast enqueued", e->comm->rank);
    B200_CHECK(cudaStreamSynchronize(e->stream));
    B200_COMM_TRACE("rank %d: broadcast done", e->comm->rank);
    conv_stem_invalidate_bank();                    // the first layer's constant-bank copy follows the arena
    return 0;
}

// From now on every rank's records carry GLOBAL image numbers (image_base + local index) and b200_detect_batch /
// b200_detect_submitted on `root` return the records of ALL ranks, in rank order; the other ranks keep returning their own.  slot_records bounds what one rank can
// contribute per batch (records beyond it are dropped and reported on stderr by the root).  root = -1 switches it off.
extern "C" int b200_comm_set_gather(network *net, int root, int image_base, int slot_records)
{
    b200_engine *e = b200_engine_of(net);
    if (!e->comm) { fprintf(stderr, "b200-darknet: b200_comm_set_gather before b200_comm_init\n"); abort(); }
    B200Comm *c = e->comm;
    B200_CHECK(cudaSetDevice(e->device));
    B200_CHECK(cudaStreamSynchronize(e->stream));
    c->gather_root = root; c->image_base = image_base;
    if (root < 0) return 0;
    if (slot_records < 1) slot_records = 1;
    if (!e->d_record_count) { B200_CHECK(cudaMalloc(&e->d_record_count, 4 * sizeof(int))); B200_CHECK(cudaMemset(e->d_record_count, 0, 4 * sizeof(int))); }
    B200_CHECK(cudaMemcpy(e->d_record_count + 1, &image_base, sizeof(int), cudaMemcpyHostToDevice));
    if (e->records_cap < slot_records) {            // a sender always ships a whole slot: the record buffer must cover it
        cudaFree(e->d_records);
        B200_CHECK(cudaMalloc(&e->d_records, (size_t)slot_records * sizeof(DetRecord)));
        e->records_cap = slot_records;
    }
    if (c->rank == root && (c->slot_records != slot_records || !c->d_slots)) {
        cudaFree(c->d_slots); cudaFree(c->d_headers);
        if (c->h_headers) cudaFreeHost(c->h_headers);
        B200_CHECK(cudaMalloc(&c->d_slots, (size_t)c->world * slot_records * sizeof(DetRecord)));
        B200_CHECK(cudaMalloc(&c->d_headers, (size_t)c->world * 4 * sizeof(int)));
        B200_CHECK(cudaMallocHost(&c->h_headers, (size_t)c->world * 4 * sizeof(int)));
    }
    c->slot_records = slot_records;
    return 0;
}

bool b200_comm_gathers(const b200_engine *e) { return e->comm && e->comm->gather_root >= 0 && e->comm->world > 1; }
bool b200_comm_is_root(const b200_engine *e) { return e->comm && e->comm->rank == e->comm->gather_root; }
int  b200_comm_image_base(const b200_engine *e) { return b200_comm_gathers(e) ? e->comm->image_base : 0; }

void b200_comm_enqueue_gather(b200_engine *e, cudaStream_t s)
{
    B200Comm *c = e->comm;
    NcclApi *a = need_nccl("detection gather");
    const size_t slot_bytes = (size_t)c->slot_records * sizeof(DetRecord);
    B200_NCCL_CHECK(a, a->GroupStart());
    if (c->rank == c->gather_root) {
        for (int r = 0; r < c->world; ++r) {
            if (r == c->rank) continue;
            B200_NCCL_CHECK(a, a->Recv(c->d_headers + 4 * r, 4, ncclInt32, r, c->comm, s));
            B200_NCCL_CHECK(a, a->Recv((unsigned char *)c->d_slots + r * slot_bytes, slot_bytes, ncclUint8, r, c->comm, s));
        }
    } else {
        B200_NCCL_CHECK(a, a->Send(e->d_record_count, 4, ncclInt32, c->gather_root, c->comm, s));
        B200_NCCL_CHECK(a, a->Send(e->d_records, slot_bytes, ncclUint8, c->gather_root, c->comm, s));
    }
    B200_NCCL_CHECK(a, a->GroupEnd());
    B200_COMM_TRACE("rank %d: gather enqueued (%zu slot bytes, root %d)", c->rank, slot_bytes, c->gather_root);
}

int b200_comm_collect_gathered(b200_engine *e, b200_det *out, int max_out, int own_count, cudaStream_t s)
{
    B200Comm *c = e->comm;
    B200_CHECK(cudaMemcpyAsync(c->h_headers, c->d_headers, (size_t)c->world * 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CHECK(cudaStreamSynchronize(s));
    int total = 0;
    for (int r = 0; r < c->world; ++r) {
        int n = r == c->rank ? own_count : c->h_headers[4 * r];
        if (n > c->slot_records && r != c->rank) {
            fprintf(stderr, "b200-darknet: rank %d produced %d records, its gather slot holds %d: the rest is dropped\n", r, n, c->slot_records);
            n = c->slot_records;
        }
        if (total + n > max_out) n = max_out - total;
        if (n > 0) {                                   // image numbers are already global: every rank's collect kernel adds its base
            const DetRecord *src = r == c->rank ? e->d_records : c->d_slots + (size_t)r * c->slot_records;
            B200_CHECK(cudaMemcpyAsync(out + total, src, (size_t)n * sizeof(DetRecord), cudaMemcpyDeviceToHost, s));
        }
        total += n;
    }
    B200_CHECK(cudaStreamSynchronize(s));
    return total;
}

void b200_comm_release(b200_engine *e)
{
    B200Comm *c = e->comm;
    if (!c) return;
    NcclApi *a = nccl_api();
    cudaFree(c->d_slots); cudaFree(c->d_headers);
    if (c->h_headers) cudaFreeHost(c->h_headers);
    if (a && c->comm) a->CommDestroy(c->comm);
    delete c;
    e->comm = nullptr;
}

extern "C" void b200_comm_destroy(network *net) { b200_comm_release(b200_engine_of(net)); }
