// libpifpaf_b200 -- image preprocessing on the GPU.
//
// Replaces, for the inference path of the reference's Predictor (predictor.py:85-102):
//   transforms.RescaleAbsolute(long_edge, fast=True)   transforms/scale.py:154-176 -> PIL.Image.resize(BILINEAR)
//   transforms.CenterPad / CenterPadTight              transforms/pad.py:15-110
// ToTensor + Normalize (transforms/__init__.py:26-33) already run inside the stem kernel (pifpaf_net_forward_u8).
//
// The reference resizes with Pillow (a third-party dependency of the reference, setup.py `pillow`; 12.2.0 in this
// image) unless OpenCV is importable.  Pillow's ImagingResample (src/libImaging/Resample.c) is restated here: a
// separable convolution, horizontal pass first into an 8-bit intermediate, then the vertical pass; coefficients are
// the antialiased triangle filter, normalised, in fixed point with 22 fractional bits (host side:
// openpifpaf_b200/preprocess.py, IEEE double arithmetic in Pillow's order), accumulated in 32-bit integers from
// 1 << 21 and shifted down -- integer arithmetic, so the resized image equals Pillow's bit for bit.
#include <algorithm>
#include <cstdint>

#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [h][w][3] -> tmp [h][tw][3]
__global__ void __launch_bounds__(256) k_resize_h(const uint8_t* __restrict__ src, int h, int w, int tw,
                                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                  uint8_t* __restrict__ tmp) {
    const long long total = (long long)h * tw;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(t % tw), y = (int)(t / tw);
        const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
        const int* k = kk + (size_t)xx * ksize;
        const uint8_t* row = src + ((size_t)y * w + xmin) * 3;
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < xmax; x++) {
            const int kv = k[x];
            s0 += (int)row[3 * x] * kv; s1 += (int)row[3 * x + 1] * kv; s2 += (int)row[3 * x + 2] * kv;
        }
        uint8_t* o = tmp + ((size_t)y * tw + xx) * 3;
        o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
    }
}

// vertical pass: tmp [h][tw][3] -> dst rows [th] x [tw] pixels at (left, top) of a canvas with row pitch dst_pitch
__global__ void __launch_bounds__(256) k_resize_v(const uint8_t* __restrict__ tmp, int tw, int th,
                                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                  uint8_t* __restrict__ dst, long long dst_pitch) {
    const long long total = (long long)th * tw;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(t % tw), yy = (int)(t / tw);
        const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
        const int* k = kk + (size_t)yy * ksize;
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < ymax; y++) {
            const uint8_t* p = tmp + ((size_t)(y + ymin) * tw + xx) * 3;
            const int kv = k[y];
            s0 += (int)p[0] * kv; s1 += (int)p[1] * kv; s2 += (int)p[2] * kv;
        }
        uint8_t* o = dst + (size_t)yy * dst_pitch + (size_t)xx * 3;
        o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
    }
}

// plain copy of an image into the canvas (no rescale in that direction pair)
__global__ void __launch_bounds__(256) k_copy_rows(const uint8_t* __restrict__ src, int h, int row_bytes,
                                                   uint8_t* __restrict__ dst, long long dst_pitch) {
    const long long total = (long long)h * row_bytes;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % row_bytes), y = (int)(t / row_bytes);
        dst[(size_t)y * dst_pitch + x] = src[(size_t)y * row_bytes + x];
    }
}

__global__ void __launch_bounds__(256) k_fill_rgb(uint8_t* __restrict__ dst, long long n_pixels, uint8_t r, uint8_t g, uint8_t b) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pixels; t += (long long)gridDim.x * blockDim.x) {
        dst[3 * t] = r; dst[3 * t + 1] = g; dst[3 * t + 2] = b;
    }
}

inline int grid_for(long long total) { return (int)std::min<long long>((total + 255) / 256, 148LL * 16); }

}  // namespace

extern "C" {

int pifpaf_image_fill_rgb(uint8_t* dst_dev, int64_t n_pixels, int32_t r, int32_t g, int32_t b, void* stream_v) {
    PIFPAF_CHECK_ARG(dst_dev != nullptr && n_pixels >= 0, "bad fill arguments");
    PIFPAF_CHECK_ARG(r >= 0 && r <= 255 && g >= 0 && g <= 255 && b >= 0 && b <= 255, "fill colour out of range");
    if (n_pixels == 0) return PIFPAF_OK;
    k_fill_rgb<<<grid_for(n_pixels), 256, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(dst_dev, n_pixels, (uint8_t)r, (uint8_t)g, (uint8_t)b);
    PIFPAF_LAUNCH_CHECK();
    return PIFPAF_OK;
}

int pifpaf_image_resize_bilinear_u8(const uint8_t* src_dev, int32_t src_h, int32_t src_w,
                                    uint8_t* dst_dev, int64_t dst_pitch_bytes, int32_t dst_h, int32_t dst_w,
                                    const int32_t* xbounds_dev, const int32_t* xkk_dev, int32_t xksize,
                                    const int32_t* ybounds_dev, const int32_t* ykk_dev, int32_t yksize,
                                    uint8_t* tmp_dev, void* stream_v) {
    PIFPAF_CHECK_ARG(src_dev != nullptr && dst_dev != nullptr, "image pointer is null");
    PIFPAF_CHECK_ARG(src_h >= 1 && src_w >= 1 && dst_h >= 1 && dst_w >= 1, "bad image size");
    PIFPAF_CHECK_ARG(dst_pitch_bytes >= (int64_t)dst_w * 3, "destination pitch too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
    const bool need_h = dst_w != src_w, need_v = dst_h != src_h;          // Resample.c: ImagingResample
    PIFPAF_CHECK_ARG(!need_h || (xbounds_dev && xkk_dev && xksize >= 1), "horizontal coefficients missing");
    PIFPAF_CHECK_ARG(!need_v || (ybounds_dev && ykk_dev && yksize >= 1), "vertical coefficients missing");
    PIFPAF_CHECK_ARG(!(need_h && need_v) || tmp_dev != nullptr, "intermediate buffer [src_h][dst_w][3] missing");
    if (need_h && need_v) {
        k_resize_h<<<grid_for((long long)src_h * dst_w), 256, 0, st>>>(src_dev, src_h, src_w, dst_w, xbounds_dev, xkk_dev, xksize, tmp_dev);
        PIFPAF_LAUNCH_CHECK();
        k_resize_v<<<grid_for((long long)dst_h * dst_w), 256, 0, st>>>(tmp_dev, dst_w, dst_h, ybounds_dev, ykk_dev, yksize, dst_dev, dst_pitch_bytes);
        PIFPAF_LAUNCH_CHECK();
    } else if (need_h) {
        // rows go straight to the canvas: the "intermediate" is the destination (pitch handled by a per-row launch
        // shape: tmp layout [h][tw][3] only matches a dense destination, so resize into tmp when the pitch differs)
        if (dst_pitch_bytes == (int64_t)dst_w * 3) {
            k_resize_h<<<grid_for((long long)src_h * dst_w), 256, 0, st>>>(src_dev, src_h, src_w, dst_w, xbounds_dev, xkk_dev, xksize, dst_dev);
            PIFPAF_LAUNCH_CHECK();
        } else {
            PIFPAF_CHECK_ARG(tmp_dev != nullptr, "intermediate buffer missing");
            k_resize_h<<<grid_for((long long)src_h * dst_w), 256, 0, st>>>(src_dev, src_h, src_w, dst_w, xbounds_dev, xkk_dev, xksize, tmp_dev);
            PIFPAF_LAUNCH_CHECK();
            k_copy_rows<<<grid_for((long long)dst_h * dst_w * 3), 256, 0, st>>>(tmp_dev, dst_h, dst_w * 3, dst_dev, dst_pitch_bytes);
            PIFPAF_LAUNCH_CHECK();
        }
    } else if (need_v) {
        k_resize_v<<<grid_for((long long)dst_h * dst_w), 256, 0, st>>>(src_dev, dst_w, dst_h, ybounds_dev, ykk_dev, yksize, dst_dev, dst_pitch_bytes);
        PIFPAF_LAUNCH_CHECK();
    } else {
        k_copy_rows<<<grid_for((long long)dst_h * dst_w * 3), 256, 0, st>>>(src_dev, dst_h, dst_w * 3, dst_dev, dst_pitch_bytes);
        PIFPAF_LAUNCH_CHECK();
    }
    return PIFPAF_OK;
}

}  // extern "C"
