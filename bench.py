#!/usr/bin/env python3
"""bench.py — YOLOv3-416 images/s through the darknet C API of this engine (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this engine (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation (oracle/_ref)

A "step" = one pass of the hot path over one batch: network forward (conv stack) + device decode +
class-wise NMS for all 64 images of the batch (configs[2] of BASELINE.json: YOLOv3 416x416 batch 64 bf16).
  value : images/s with the batch already resident in HBM (b200_detect_batch with input=NULL), CUDA-event
          timed on the engine's stream, max over ranks; per-GPU batch is fixed => weak scaling.
  e2e   : the same metric through the C-ABI serving loop with HOST (pinned) fp32 NCHW batches: every timed step
          H2D-copies one 133 MB batch (b200_submit_batch, double-buffered so it overlaps the previous step's compute)
          and D2H-reads its kept detections (b200_detect_submitted).
  roofline     : tensor-pipe roofline of the dominant kernel family (conv_tc): darknet's own BFLOPs formula
                 x images / CUDA-event time of those launches, against MEASURED_PEAKS.json.
  cpu_baseline : the unmodified reference CPU build (oracle/_ref, GPU=0 OPENMP=1) timed on this box's host
                 cores on a bounded sample of the same workload (rank 0, N=1 only).
PyTorch is used for torch.distributed (NCCL weight broadcast, barriers), pinned host memory and CUDA events
on the engine's stream; every kernel in the timed region is this repo's own.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

MODEL, SIZE, BATCH = "yolov3", 416, 64
THRESH, NMS = 0.5, 0.45
FLOP_PER_IMAGE = 65_864_075_264            # reference `darknet ops` on the derived yolov3.cfg
WORK = os.environ.get("B200_WORKDIR", "/tmp/b200_bench")


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), hbm=p.get("hbm_gbs"), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.t0 = self.t1 = None                      # samples are kept between window_start() and window_end()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active,power.draw,power.limit")
        # one streaming nvidia-smi (a line every 20 ms): the timed region is only ~0.1 s long, so polling a fresh process per
        # sample would see it once
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        try:
            for line in proc.stdout:
                line = line.strip()
                if line and self.t0 is not None and (self.t1 is None or time.time() <= self.t1 + 0.02):
                    self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        finally:
            proc.terminate()
            try:
                proc.wait(timeout=2)
            except Exception:
                proc.kill()

    def window_start(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()
        self.stop_flag = True

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j].lower().startswith("active") for r in self.rows)]
        # every other bit of the event-reason mask seen under load (nvml: 0x1 idle, 0x2 application clocks, 0x10 sync boost,
        # 0x80 power brake, 0x100 display clocks), so that clocks below max never come without their reason
        other = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x10: "sync_boost", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        mask = 0
        for r in self.rows:
            try:
                mask |= int(r[6], 16) if len(r) > 6 else 0
            except ValueError:
                pass
        reasons += [n for bit, n in other.items() if mask & bit]

        def med(col):
            v = []
            for r in self.rows:
                try:
                    v.append(float(r[col]))
                except (ValueError, IndexError):
                    pass
            return sorted(v)[len(v) // 2] if v else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "sm_min_mhz": sm[0] if sm else None, "reasons": reasons, "samples": len(self.rows),
                "power_w": med(7), "power_limit_w": med(8)}


def prepare_files(batch):
    from yolo_tensorflow_b200 import synth
    os.makedirs(WORK, exist_ok=True)
    cfg = synth.make_cfg(MODEL, WORK, batch=batch, width=SIZE, height=SIZE)
    wpath = os.path.join(WORK, f"{MODEL}_seed0_damped.weights")
    if not os.path.exists(wpath):
        tmp = wpath + f".{os.getpid()}.tmp"
        synth.write_weights(cfg, tmp, seed=0, damp_heads=True)
        os.replace(tmp, wpath)
    return cfg, wpath


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle/_ref), host cores only
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, images_per_step=1):
    """times network_predict + get_network_boxes + do_nms_sort (detector.c:596-603 idiom) per image"""
    from yolo_tensorflow_b200 import synth
    from oracle import ref_darknet as R
    cfg, wpath = prepare_files(1)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host thread it can
    os.environ["OMP_NUM_THREADS"] = os.environ.get("B200_REF_THREADS", str(cores))
    kind = "reference"
    if R.available():
        net = R.RefNet(cfg, wpath)

        def one(img):
            net.predict(img)
            dets, n = net.boxes(0, SIZE, SIZE, THRESH)
            net.nms_sort(dets, n, NMS)
            net.free_dets(dets, n)
    else:                                   # the reference .so did not travel: fall back to the numpy restatement
        from oracle import np_darknet as P
        kind = "port"
        net = P.Net(cfg, wpath)

        def one(img):
            outs = net.forward(img)
            b, o, p, _ = P.get_network_boxes(net, outs, 0, SIZE, SIZE, THRESH)
            P.do_nms_sort(b, o, p, NMS)
    imgs = synth.make_images(max(1, images_per_step), 3, SIZE, SIZE, 1002)
    for _ in range(warmup):
        one(imgs[:1])
    t0 = time.perf_counter()
    for _ in range(steps):
        for j in range(images_per_step):
            one(imgs[j:j + 1])
    dt = time.perf_counter() - t0
    return dict(value=steps * images_per_step / dt, seconds=dt, kind=kind, cores=cores,
                sample=f"{steps * images_per_step} image(s) of YOLOv3-416 (predict + get_network_boxes + do_nms_sort), batch 1 as the reference runs it")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))
    r = cpu_reference_run(steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": "YOLOv3-416 images/s (conv+decode+NMS)", "value": r["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * r["seconds"] / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"YOLOv3 {SIZE}x{SIZE} batch {BATCH} (reference CPU arm: 1 image per step, batch 1)",
                       "thresh": THRESH, "nms": NMS},
            "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# this engine
# ---------------------------------------------------------------------------------------------------
class _DevPtr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main_engine(args):
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from yolo_tensorflow_b200 import synth, darknet as dn

    if rank == 0:
        cfg, wpath = prepare_files(BATCH)
    if world > 1:
        dist.barrier()
    if rank != 0:
        cfg = os.path.join(WORK, f"{MODEL}_b{BATCH}_{SIZE}x{SIZE}.cfg")
        if not os.path.exists(cfg):
            from yolo_tensorflow_b200 import synth as s2
            cfg = s2.make_cfg(MODEL, WORK, batch=BATCH, width=SIZE, height=SIZE)
    dn.set_gpu(local)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2); os.dup2(devnull, 2)                 # the layer table goes to stderr like the reference's
    try:
        # weights: rank 0 reads the .weights file, folds/repacks/uploads; the other ranks only parse the cfg and
        # receive the finished parameter arena with ONE NCCL broadcast over NVLink
        net = dn.Network(cfg, wpath if rank == 0 else None, precision=dn.PREC_BF16)
    finally:
        os.dup2(saved, 2); os.close(saved); os.close(devnull)
    if world > 1:
        ptr, nbytes = net.weights_arena()
        arena = torch.as_tensor(_DevPtr(ptr, nbytes), device=torch.device("cuda", local))
        dist.broadcast(arena, src=0)
        torch.cuda.synchronize()
    net.set_head_sync(0)

    # per-rank image shard: a different seeded batch on every rank
    x_host = torch.from_numpy(synth.make_images(BATCH, 3, SIZE, SIZE, 1002 + rank)).pin_memory()
    stream = torch.cuda.ExternalStream(net.stream_ptr(), device=torch.device("cuda", local))
    in_ptr, in_bytes = net.input_device_ptr(), BATCH * 3 * SIZE * SIZE * 4
    d_in = torch.as_tensor(_DevPtr(in_ptr, in_bytes), device=torch.device("cuda", local))
    d_in.copy_(x_host.view(torch.uint8).reshape(-1))
    torch.cuda.synchronize()

    max_out = 1 << 20
    out = (dn.B200_DET * max_out)()
    counts = (dn.c_int * BATCH)()

    # value: the batch stays resident in HBM; every step is a full forward + decode + NMS + collect + read-back of its records.
    # The steps run through the same one-deep pipelined call as the serving loop (the next step's forward is enqueued before
    # this step's records are awaited), with B200_INPUT_RESIDENT instead of a host batch.
    RESIDENT = ctypes.c_void_p(1)
    resident_state = {"primed": False}

    def step_resident():
        if not resident_state["primed"]:
            dn.lib.b200_submit_batch(net.ptr, RESIDENT)
            resident_state["primed"] = True
        return dn.lib.b200_detect_submitted(net.ptr, RESIDENT, SIZE, SIZE, THRESH, NMS, 1, out, max_out, counts)

    def drain_resident():
        if resident_state["primed"]:
            dn.lib.b200_detect_submitted(net.ptr, None, SIZE, SIZE, THRESH, NMS, 1, out, max_out, counts)
            resident_state["primed"] = False

    # e2e: the serving loop a user of the C API runs — every step H2D-copies its own pinned host batch and D2H-reads its
    # detections; the copy of batch k+1 is submitted before batch k is computed (double-buffered device input)
    x_host2 = torch.from_numpy(synth.make_images(BATCH, 3, SIZE, SIZE, 2002 + rank)).pin_memory()
    host_batches = [x_host, x_host2]
    e2e_state = {"k": 0}

    def step_e2e():
        k = e2e_state["k"]
        e2e_state["k"] = k + 1
        # results of batch k; the H2D of batch k+1 is started inside the call, right after batch k became current
        return dn.lib.b200_detect_submitted(net.ptr, host_batches[(k + 1) % 2].data_ptr(), SIZE, SIZE, THRESH, NMS, 1, out, max_out, counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = 0
        for _ in range(steps):
            n = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                               # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(max(args.warmup, 3)):
        step_resident()
    if sampler:
        sampler.window_start()
    launches0 = dn.lib.b200_launch_count()
    ms, nrec = timed(step_resident, args.steps)
    launches = dn.lib.b200_launch_count() - launches0
    if sampler:
        sampler.window_end()
    cand = float(np.mean(list(counts)))
    value = world * BATCH * args.steps / (ms / 1000.0)
    drain_resident()                                  # the forward pass pre-enqueued for a step that will not come

    dn.lib.b200_submit_batch(net.ptr, host_batches[0].data_ptr())      # prime the pipeline: every timed step still copies one batch
    for _ in range(2):
        step_e2e()
    ms_e2e, nrec_e2e = timed(step_e2e, args.steps)
    dn.lib.b200_detect_submitted(net.ptr, None, SIZE, SIZE, THRESH, NMS, 1, out, max_out, counts)     # drain the last submitted batch
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1000.0)

    # detections gathered once at the end (variable length): counts first, then the records
    total_records = nrec
    if world > 1:
        t = torch.tensor([nrec], device="cuda")
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        total_records = int(sum(int(g.item()) for g in gathered))

    line = None
    if rank == 0:
        pk = peaks()
        # roofline of the dominant kernel family, measured live with CUDA events between layers
        per_layer = net.profile_layers(3)
        is_tc = [net.kernel(i).startswith("conv_tc") for i in range(net.n)]      # "conv_tc" and "conv_tc+shortcut" launches
        conv_ms = sum(float(per_layer[i]) for i in range(net.n) if is_tc[i])
        conv_flops = sum(2.0 * L["n"] * L["size"] ** 2 * L["c"] * L["out_h"] * L["out_w"] * BATCH
                         for i, L in enumerate(net.layers) if is_tc[i])
        n_conv = sum(is_tc)
        by_kernel = {}
        for i in range(net.n):
            by_kernel[net.kernel(i)] = by_kernel.get(net.kernel(i), 0.0) + float(per_layer[i])
        achieved = conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
        traffic = None                      # DRAM bytes moved by the conv_tc family per step, from the committed ncu capture
        tpath = os.path.join(REPO, "profiles", "r1_conv_tc_dram_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_total")
        roofline = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, %d launches/step)" % n_conv,
                    "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                    "traffic": traffic,
                    "traffic_note": "dram__bytes_read+write summed over the conv_tc launches of one step (profiles/r1_conv_tc_dram_traffic.json); "
                                    "algorithmic activation+weight bytes of those layers at bf16: %.2f GB" % (
                                        sum((L["c"] * L["h"] * L["w"] + L["out_c"] * L["out_h"] * L["out_w"]) * 2.0 * BATCH + L["nweights"] * 2.0
                                            for i, L in enumerate(net.layers) if is_tc[i]) / 1e9),
                    "peak_source": pk["source"], "conv_ms_per_step": conv_ms,
                    "ms_per_step_by_kernel": {k: round(v, 4) for k, v in by_kernel.items()},
                    "whole_step_tflops": FLOP_PER_IMAGE * BATCH * args.steps / (ms / 1000.0) / 1e12 / max(world, 1)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(4, 1)
            cpu = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        line = {"metric": "YOLOv3-416 images/s (conv+decode+NMS)", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"YOLOv3 {SIZE}x{SIZE} batch {BATCH} per GPU, bf16 activations, fp32 accumulate; step = forward + decode + NMS",
                           "thresh": THRESH, "nms": NMS, "mean_candidates_per_image": cand, "kept_records_per_step": total_records,
                           "l2": "working set per step (133 MB input + >5 GB activations) exceeds the 126 MB L2; no explicit flush",
                           "pipelining": "value and e2e both use b200_detect_submitted: step k+1's forward is enqueued before step k's records are awaited (value: B200_INPUT_RESIDENT, e2e: pinned host batches)",
                           "weights": "seed-0 synthetic, damped heads (yolo_tensorflow_b200/synth.py)",
                           "multi_gpu": "image-sharded, weights NCCL-broadcast once, no per-layer collective"},
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": in_bytes,
                        "d2h_bytes_per_step": int(nrec_e2e) * 36 + 4 * BATCH + 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "clocks": sampler.summary() if sampler else None}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    net.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_engine(a)
