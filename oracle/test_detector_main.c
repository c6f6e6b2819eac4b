/* TEST INFRASTRUCTURE.  A main() around the reference's own `test_detector` (examples/detector.c:562-627): the function's
 * text is extracted from /root/reference at build time into oracle/_ref/test_detector_body.inc (git-ignored, never
 * committed) and compiled UNCHANGED twice by oracle/Makefile — against the reference header + library
 * (oracle/_ref/test_detector_ref) and against this repository's public include/darknet.h + libdarknet.so
 * (oracle/_ref/test_detector_b200).  tests/test_drop_in_driver.py runs both on the same files and compares what they print.
 * usage: test_detector_* <datacfg> <cfg> <weights> <image> <thresh> <outfile-prefix> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "darknet.h"

#include "test_detector_body.inc"

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s datacfg cfg weights image thresh outfile\n", argv[0]); return 2; }
    test_detector(argv[1], argv[2], argv[3], argv[4], atof(argv[5]), .5, argv[6], 0);
    return 0;
}
