// Batch assembly on the device: gather -> RandomResizedCrop -> (grey -> RGB) -> RandomHorizontalFlip in one pass.
//
// Replaces, for a whole per-rank batch at once, what the reference does per sample on DataLoader workers:
// direct_dataset.__getitem__ (main_direct.py:200-204) = numpy row -> torchvision RandomResizedCrop(size, scale=(0.5,
// 1.0)) -> Lambda(repeat to 3 channels) -> RandomHorizontalFlip (main_direct.py:158-169), i.e. crop, an ATen bilinear
// resize (align_corners=False), a repeat and a flip copy per image, then collation and a 154 MB host -> device
// copy per 256-image step.  Here the image set lives in HBM (a B200 holds ~300 000 fp32 224x224 images beside the
// model) and a batch is one kernel: every output pixel reads its four taps straight from the stored NCHW image
// of its sample (neighbouring outputs share taps through L1/L2) and is written once, in the memory format the
// QAT step runs in.  The random draws (boxes, flips) stay on the host -- ood_dfq_b200/augment.py restates
// torchvision's get_params draw for draw -- and arrive as two small device arrays.
//
// Roofline: HBM.  Algorithmic bytes per sample: 4 * C_in * crop_h * crop_w read + 4 * C_out * OH * OW written.
#include "augment_core.h"
#include "common.cuh"

namespace oodfq {

template <int C_IN, int C_OUT, int PX>
__global__ void __launch_bounds__(256)
crop_resize_flip_kernel(const AugArgs a, int total) {
    const int groups = (total + PX - 1) / PX;                 // total <= 2^31 - 4 (checked by the entry point)
    for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups;
         gi += (long long)gridDim.x * blockDim.x)
        aug_group<C_IN, C_OUT, PX>(a, (int)gi * PX, total);
}

// backward: one thread per output pixel, atomic scatter onto the (zero-initialised or accumulating) image gradient
template <int C_IN, int C_OUT>
__global__ void __launch_bounds__(256)
crop_resize_flip_bwd_kernel(const AugArgs a, int total) {
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x)
        aug_pixel_backward<C_IN, C_OUT>(a, (int)pix);
}

template <int C_IN, int C_OUT>
static void launch_aug_bwd(const AugArgs& a, int total, cudaStream_t st) {
    static const int per_sm = resident_ctas(crop_resize_flip_bwd_kernel<C_IN, C_OUT>, 256);
    const long long want = ((long long)total + 255) / 256, cap = (long long)kNumSM * per_sm * 4;
    crop_resize_flip_bwd_kernel<C_IN, C_OUT><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(a, total);
}

template <int C_IN, int C_OUT, int PX>
static void launch_aug(const AugArgs& a, int total, cudaStream_t st) {
    static const int per_sm = resident_ctas(crop_resize_flip_kernel<C_IN, C_OUT, PX>, 256);
    const long long groups = (total + PX - 1) / PX, want = (groups + 255) / 256, cap = (long long)kNumSM * per_sm * 4;
    crop_resize_flip_kernel<C_IN, C_OUT, PX><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(a, total);
}

}  // namespace oodfq

using namespace oodfq;

static int aug_check(const char* what, long long n_images, int C_in, int H, int W, int N, int C_out, int out_h, int out_w) {
    if (n_images <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0)
        return fail(OODFQ_EINVAL, "%s: empty image set or output (M=%lld, %dx%d -> %dx%d)", what, n_images, H, W, out_h, out_w);
    if (!((C_in == 1 && (C_out == 1 || C_out == 3)) || (C_in == 3 && C_out == 3)))
        return fail(OODFQ_EINVAL, "%s: channels %d -> %d (supported: 1->1, 1->3, 3->3)", what, C_in, C_out);
    if (H > kAugMaxSide || W > kAugMaxSide || out_h > kAugMaxSide || out_w > kAugMaxSide)
        return fail(OODFQ_EINVAL, "%s: sides above %d are not supported", what, kAugMaxSide);
    if ((long long)N * out_h * out_w > 0x7fffffffLL - 4)
        return fail(OODFQ_EINVAL, "%s: more than 2^31 output pixels in one call; split the batch", what);
    return OODFQ_OK;
}

static AugArgs aug_args(const float* images, long long n_images, int C_in, int H, int W, const long long* index,
                        const int* boxes, const unsigned char* flips, float* out, const float* grad_out, int N, int out_h,
                        int out_w, int flags) {
    AugArgs a;
    a.images = images; a.index = index; a.boxes = boxes; a.flips = flips; a.out = out; a.grad_out = grad_out;
    a.g.M = n_images; a.g.C_in = C_in; a.g.H = H; a.g.W = W; a.g.N = N; a.g.OH = out_h; a.g.OW = out_w;
    a.g.nhwc = (flags & OODFQ_BN_NHWC) ? 1 : 0;
    a.g.src_nhwc = (flags & OODFQ_AUG_SRC_NHWC) ? 1 : 0;
    return a;
}

extern "C" int oodfq_crop_resize_flip(const float* images, long long n_images, int C_in, int H, int W,
                                      const long long* index, const int* boxes, const unsigned char* flips,
                                      float* out, int N, int C_out, int out_h, int out_w, int flags,
                                      oodfq_stream_t stream) {
    if (N < 0) return fail(OODFQ_EINVAL, "crop_resize_flip: N=%d", N);
    if (N == 0) return OODFQ_OK;
    if (!images || !index || !boxes || !flips || !out) return fail(OODFQ_EINVAL, "crop_resize_flip: null pointer");
    const int rc = aug_check("crop_resize_flip", n_images, C_in, H, W, N, C_out, out_h, out_w);
    if (rc != OODFQ_OK) return rc;
    const AugArgs a = aug_args(images, n_images, C_in, H, W, index, boxes, flips, out, nullptr, N, out_h, out_w, flags);
    const int total = N * out_h * out_w;
    cudaStream_t st = (cudaStream_t)stream;
    // channels_last: four pixels per thread so that the stores are 128-bit; NCHW (and misaligned outputs): one
    // pixel per thread, each channel plane written by consecutive lanes
    const bool wide = a.g.nhwc && aligned16(out);
    if (C_in == 3) {
        if (wide) launch_aug<3, 3, 4>(a, total, st); else launch_aug<3, 3, 1>(a, total, st);
    } else if (C_out == 3) {
        if (wide) launch_aug<1, 3, 4>(a, total, st); else launch_aug<1, 3, 1>(a, total, st);
    } else {
        if (wide) launch_aug<1, 1, 4>(a, total, st); else launch_aug<1, 1, 1>(a, total, st);
    }
    count_launch();
    return check_launch("crop_resize_flip");
}

extern "C" int oodfq_crop_resize_flip_backward(const float* grad_out, float* grad_images, long long n_images, int C_in,
                                               int H, int W, const long long* index, const int* boxes,
                                               const unsigned char* flips, int N, int C_out, int out_h, int out_w,
                                               int flags, oodfq_stream_t stream) {
    if (N < 0) return fail(OODFQ_EINVAL, "crop_resize_flip_backward: N=%d", N);
    if (N == 0) return OODFQ_OK;
    if (!grad_out || !grad_images || !index || !boxes || !flips)
        return fail(OODFQ_EINVAL, "crop_resize_flip_backward: null pointer");
    const int rc = aug_check("crop_resize_flip_backward", n_images, C_in, H, W, N, C_out, out_h, out_w);
    if (rc != OODFQ_OK) return rc;
    const AugArgs a = aug_args(nullptr, n_images, C_in, H, W, index, boxes, flips, grad_images, grad_out, N, out_h, out_w,
                               flags);
    const int total = N * out_h * out_w;
    cudaStream_t st = (cudaStream_t)stream;
    if (C_in == 3) launch_aug_bwd<3, 3>(a, total, st);
    else if (C_out == 3) launch_aug_bwd<1, 3>(a, total, st);
    else launch_aug_bwd<1, 1>(a, total, st);
    count_launch();
    return check_launch("crop_resize_flip_backward");
}
