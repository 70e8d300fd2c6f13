// Host build of the product's per-pixel augmentation body (ood_dfq_b200/csrc/augment_core.h) for CPU tests.
// TEST SCAFFOLDING: compiled by tests/test_augment_cpu.py with g++, never shipped, never reachable from the
// package.  It lets the index arithmetic, the tap weights, the flip, the grey -> RGB repeat, both layouts and the
// gradient scatter of the sm_100a kernels be checked against the oracle where there is no GPU; the launch
// geometry of augment.cu is what remains for the -m gpu test.
#include "../../ood_dfq_b200/csrc/augment_core.h"

using namespace oodfq;

template <int C_IN, int C_OUT, int PX>
static void run(const AugArgs& a, int total) {
    for (int first = 0; first < total; first += PX) aug_group<C_IN, C_OUT, PX>(a, first, total);
}

template <int C_IN, int C_OUT>
static void run_bwd(const AugArgs& a, int total) {
    for (int pix = 0; pix < total; ++pix) aug_pixel_backward<C_IN, C_OUT>(a, pix);
}

static AugArgs make(const float* images, long long M, int C_in, int H, int W, const long long* index, const int* boxes,
                    const unsigned char* flips, float* out, const float* grad_out, int N, int OH, int OW, int nhwc,
                    int src_nhwc) {
    AugArgs a;
    a.images = images; a.index = index; a.boxes = boxes; a.flips = flips; a.out = out; a.grad_out = grad_out;
    a.g.M = M; a.g.C_in = C_in; a.g.H = H; a.g.W = W; a.g.N = N; a.g.OH = OH; a.g.OW = OW; a.g.nhwc = nhwc;
    a.g.src_nhwc = src_nhwc;
    return a;
}

extern "C" int augment_host(const float* images, long long M, int C_in, int H, int W, const long long* index,
                            const int* boxes, const unsigned char* flips, float* out, int N, int C_out, int OH,
                            int OW, int nhwc, int px, int src_nhwc) {
    const AugArgs a = make(images, M, C_in, H, W, index, boxes, flips, out, nullptr, N, OH, OW, nhwc, src_nhwc);
    const int total = N * OH * OW;
    if (C_in == 3 && C_out == 3) { if (px == 4) run<3, 3, 4>(a, total); else run<3, 3, 1>(a, total); return 0; }
    if (C_in == 1 && C_out == 3) { if (px == 4) run<1, 3, 4>(a, total); else run<1, 3, 1>(a, total); return 0; }
    if (C_in == 1 && C_out == 1) { if (px == 4) run<1, 1, 4>(a, total); else run<1, 1, 1>(a, total); return 0; }
    return -1;
}

// grad_images must be zero-filled (or hold a gradient to accumulate into) by the caller, as for the device entry point
extern "C" int augment_host_backward(const float* grad_out, float* grad_images, long long M, int C_in, int H, int W,
                                     const long long* index, const int* boxes, const unsigned char* flips, int N,
                                     int C_out, int OH, int OW, int nhwc, int src_nhwc) {
    const AugArgs a = make(nullptr, M, C_in, H, W, index, boxes, flips, grad_images, grad_out, N, OH, OW, nhwc, src_nhwc);
    const int total = N * OH * OW;
    if (C_in == 3 && C_out == 3) { run_bwd<3, 3>(a, total); return 0; }
    if (C_in == 1 && C_out == 3) { run_bwd<1, 3>(a, total); return 0; }
    if (C_in == 1 && C_out == 1) { run_bwd<1, 1>(a, total); return 0; }
    return -1;
}
