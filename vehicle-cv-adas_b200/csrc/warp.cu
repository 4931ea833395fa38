// warp.cu -- cv2.warpPerspective(img, M, dsize, flags=INTER_LINEAR) with the default constant (black) border, bit-exact, for a batch of
// BGR u8 frames: PerspectiveTransformation.transformToBirdView / transformToFrontalView (perspectiveTransformation.py:90-117) on
// the device ("cv2.warpPerspective, a 2.76 MB gather kernel").
//
// OpenCV's algorithm (imgwarp.cpp, WarpPerspectiveInvoker + remapBilinear), restated:
//   * the map is the INVERSE of M (cv::invert of a 3x3 double matrix = the closed-form adjugate / determinant, same operation order
//     here, on the host);
//   * destination pixels are processed in blocks bw0 wide (bw0 = min(1024 / min(16, rows), cols), i.e. 64 for frames >= 64 wide):
//       X0 = M0*x_block + M1*y + M2,  W = W0 + M6*x1,  W = W ? 32 / W : 0,  X = cvRound(clamp((X0 + M0*x1) * W, INT_MIN, INT_MAX))
//     in float64 (round half to even), likewise Y: coordinates in 1/32 pixel;
//   * source pixel (X >> 5, Y >> 5) saturated to int16, bilinear weights from the 5 fractional bits: the 2-D table is
//     round(wy * wx * 32768) of float32 products that are exact, i.e. (32 - ay or ay) * (32 - ax or ax) * 32 -- always summing to 32768;
//   * result = (sum of the four taps * weight + 16384) >> 15, taps outside the source count as 0 (BORDER_CONSTANT, value 0).
// Built with -fmad=false; the float64 steps use _rn intrinsics in OpenCV's operation order.
#include "common.h"
#include "../../include/adas_b200.h"
#include <limits.h>
#include <vector>

namespace adas {

__global__ void warp_perspective_kernel(const uint8_t* __restrict__ src, int B, int H, int W, const double* __restrict__ Minv, uint8_t* __restrict__ dst,
                                        int oh, int ow, int bw0) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= ow) return;
    const double* M = Minv + (size_t)b * 9;
    const int xb = (x / bw0) * bw0, x1 = x - xb;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], (double)xb), __dmul_rn(M[1], (double)y)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], (double)xb), __dmul_rn(M[4], (double)y)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], (double)xb), __dmul_rn(M[7], (double)y)), M[8]);
    double Wv = __dadd_rn(W0, __dmul_rn(M[6], (double)x1));
    Wv = Wv != 0.0 ? __ddiv_rn(32.0, Wv) : 0.0;
    const double fX = fmax((double)INT_MIN, fmin((double)INT_MAX, __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], (double)x1)), Wv)));
    const double fY = fmax((double)INT_MIN, fmin((double)INT_MAX, __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], (double)x1)), Wv)));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int ax = X & 31, ay = Y & 31;
    const int w00 = (32 - ay) * (32 - ax), w01 = (32 - ay) * ax, w10 = ay * (32 - ax), w11 = ay * ax;      // x 32 = OpenCV's int16 table
    const uint8_t* S = src + (size_t)b * H * W * 3;
    const bool y0ok = sy >= 0 && sy < H, y1ok = sy + 1 >= 0 && sy + 1 < H, x0ok = sx >= 0 && sx < W, x1ok = sx + 1 >= 0 && sx + 1 < W;
    uint8_t* o = dst + (((size_t)b * oh + y) * ow + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int p00 = (y0ok && x0ok) ? S[((size_t)sy * W + sx) * 3 + c] : 0;
        const int p01 = (y0ok && x1ok) ? S[((size_t)sy * W + sx + 1) * 3 + c] : 0;
        const int p10 = (y1ok && x0ok) ? S[((size_t)(sy + 1) * W + sx) * 3 + c] : 0;
        const int p11 = (y1ok && x1ok) ? S[((size_t)(sy + 1) * W + sx + 1) * 3 + c] : 0;
        const int acc = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11) * 32;
        o[c] = (uint8_t)min(255, max(0, (acc + (1 << 14)) >> 15));
    }
}

// cv::invert for a 3x3 float64 matrix (matrix_decomp / lapack.cpp: closed form with the determinant expanded along the first row)
static bool invert3x3(const double* s, double* t) {
    auto S = [&](int r, int c) { return s[r * 3 + c]; };
    double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) + S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    if (d == 0.0) return false;
    d = 1.0 / d;
    t[0] = (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d;
    t[1] = (S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d;
    t[2] = (S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d;
    t[3] = (S(1, 2) * S(2, 0) - S(1, 0) * S(2, 2)) * d;
    t[4] = (S(0, 0) * S(2, 2) - S(0, 2) * S(2, 0)) * d;
    t[5] = (S(0, 2) * S(1, 0) - S(0, 0) * S(1, 2)) * d;
    t[6] = (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)) * d;
    t[7] = (S(0, 1) * S(2, 0) - S(0, 0) * S(2, 1)) * d;
    t[8] = (S(0, 0) * S(1, 1) - S(0, 1) * S(1, 0)) * d;
    return true;
}

// frames already on the device; M: host, forward matrices [B,9]; d_Minv: device scratch [B,9]; dst: device [B,oh,ow,3]
int launch_warp_perspective(const uint8_t* d_src, int B, int H, int W, const double* M_host, double* d_Minv, uint8_t* d_dst, int oh, int ow, cudaStream_t st) {
    ADAS_CHECK(B >= 1 && B <= 1024 && H >= 1 && W >= 1 && oh >= 1 && ow >= 1 && oh <= 65535, "warp_perspective: bad geometry %dx%d -> %dx%d (batch %d)", H, W, oh, ow, B);
    double inv[9];
    std::vector<double> all((size_t)B * 9);
    for (int b = 0; b < B; ++b) {
        ADAS_CHECK(invert3x3(M_host + (size_t)b * 9, inv), "warp_perspective: matrix %d is singular", b);
        for (int k = 0; k < 9; ++k) all[(size_t)b * 9 + k] = inv[k];
    }
    ADAS_CUDA(cudaMemcpyAsync(d_Minv, all.data(), all.size() * 8, cudaMemcpyHostToDevice, st));
    ADAS_CUDA(cudaStreamSynchronize(st));          // `all` is a pageable temporary
    const int bh0 = oh < 16 ? oh : 16;
    int bw0 = 1024 / bh0;
    if (bw0 > ow) bw0 = ow;
    dim3 grid((ow + 127) / 128, oh, B);
    warp_perspective_kernel<<<grid, 128, 0, st>>>(d_src, B, H, W, d_Minv, d_dst, oh, ow, bw0);
    count_launch();
    ADAS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace adas

extern "C" int adas_warp_perspective(int device, const uint8_t* frames_host, int batch, int H, int W, const double* M, int out_h, int out_w,
                                     uint8_t* out_host) {
    using namespace adas;
    ADAS_CHECK(frames_host && M && out_host && batch >= 1, "adas_warp_perspective: null argument");
    ADAS_CUDA(cudaSetDevice(device));
    uint8_t *d_s = nullptr, *d_d = nullptr; double* d_M = nullptr;
    const size_t sb = (size_t)batch * H * W * 3, db = (size_t)batch * out_h * out_w * 3;
    int rc = 0;
    auto fail = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && !rc) { set_error("adas_warp_perspective %s: %s", what, cudaGetErrorString(e)); rc = 1; } };
    fail(cudaMalloc(&d_s, sb), "alloc"); fail(cudaMalloc(&d_d, db), "alloc"); fail(cudaMalloc(&d_M, (size_t)batch * 72), "alloc");
    if (!rc) fail(cudaMemcpy(d_s, frames_host, sb, cudaMemcpyHostToDevice), "H2D");
    if (!rc) rc = launch_warp_perspective(d_s, batch, H, W, M, d_M, d_d, out_h, out_w, 0);
    if (!rc) fail(cudaMemcpy(out_host, d_d, db, cudaMemcpyDeviceToHost), "D2H");
    cudaFree(d_s); cudaFree(d_d); cudaFree(d_M);
    return rc;
}
