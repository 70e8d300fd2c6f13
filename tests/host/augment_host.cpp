// Host build of the product's per-pixel augmentation body (ood_dfq_b200/csrc/augment_core.h) for CPU tests.
// TEST SCAFFOLDING: compiled by tests/test_augment_cpu.py with g++, never shipped, never reachable from the
// package.  It lets the index arithmetic, the tap weights, the flip, the grey -> RGB repeat and both output
// layouts of the sm_100a kernel be checked against the oracle where there is no GPU; the launch geometry of
// augment.cu is what remains for the -m gpu test.
#include "../../ood_dfq_b200/csrc/augment_core.h"

using namespace oodfq;

template <int C_IN, int C_OUT, int PX>
static void run(const AugArgs& a, int total) {
    for (int first = 0; first < total; first += PX) aug_group<C_IN, C_OUT, PX>(a, first, total);
}

extern "C" int augment_host(const float* images, long long M, int C_in, int H, int W, const long long* index,
                            const int* boxes, const unsigned char* flips, float* out, int N, int C_out, int OH,
                            int OW, int nhwc, int px) {
    AugArgs a;
    a.images = images; a.index = index; a.boxes = boxes; a.flips = flips; a.out = out;
    a.g.M = M; a.g.C_in = C_in; a.g.H = H; a.g.W = W; a.g.N = N; a.g.OH = OH; a.g.OW = OW; a.g.nhwc = nhwc;
    const int total = N * OH * OW;
    if (C_in == 3 && C_out == 3) { if (px == 4) run<3, 3, 4>(a, total); else run<3, 3, 1>(a, total); return 0; }
    if (C_in == 1 && C_out == 3) { if (px == 4) run<1, 3, 4>(a, total); else run<1, 3, 1>(a, total); return 0; }
    if (C_in == 1 && C_out == 1) { if (px == 4) run<1, 1, 4>(a, total); else run<1, 1, 1>(a, total); return 0; }
    return -1;
}
