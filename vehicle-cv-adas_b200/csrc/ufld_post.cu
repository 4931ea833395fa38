// ufld_post.cu -- UFLDv2 row/column-anchor decode on the device (compiled with -fmad=false).
//
// Replaces UltrafastLaneDetectorV2.__process_output + _softmax
//   TrafficLaneDetector/ufldDetector/ultrafastLaneDetectorV2.py:114-181, :15-19
//   - max_indices = loc.argmax(grid axis) (first maximum), valid = exist.argmax(axis=1) (tie -> 0)   :127-134
//   - row lanes {1,2}: lane kept iff sum(valid) > num_cls_row/2;  col lanes {0,3}: > num_cls_col/4    :148,166
//   - per valid anchor: float32 softmax over the <=3 bins around the argmax, float64 expectation
//     (sum p*ind + 0.5)/(num_grid-1)*image_size, int() truncation                                    :151-154,169-172
//   - lane order left-side(col 0), left-ego(row 1), right-ego(row 2), right-side(col 3); detected iff > 2 points
//
// One CTA per frame.  Thread t < ncr*nl scans the row head for (anchor k = t/nl, lane i = t%nl) so that
// consecutive threads read consecutive floats of loc_row[g][k][i]; the column head likewise.
#include "common.h"

namespace adas {

static constexpr int UFLD_THREADS = 384;
static constexpr int UFLD_MAX_ANCH = 128;

__global__ void __launch_bounds__(UFLD_THREADS)
ufld_post_kernel(const float* __restrict__ heads, int ld, UfldDims d, int img_w, int img_h,
                 const double* __restrict__ row_anchor, const double* __restrict__ col_anchor, int32_t* __restrict__ pts,
                 int32_t* __restrict__ npts, uint8_t* __restrict__ status, double* __restrict__ coords, int max_pts) {
    __shared__ int s_valid[2][4][UFLD_MAX_ANCH];      // [row/col][lane][anchor]
    __shared__ double s_coord[2][4][UFLD_MAX_ANCH];   // expectation coordinate (float64, pre-truncation)
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float* loc_row = heads + (size_t)b * ld;
    const float* loc_col = loc_row + (size_t)d.ngr * d.ncr * d.nl;
    const float* ex_row = loc_col + (size_t)d.ngc * d.ncc * d.nl;
    const float* ex_col = ex_row + (size_t)2 * d.ncr * d.nl;

    for (int part = 0; part < 2; ++part) {
        const int ng = part == 0 ? d.ngr : d.ngc;
        const int ncls = part == 0 ? d.ncr : d.ncc;
        const float* loc = part == 0 ? loc_row : loc_col;
        const float* ex = part == 0 ? ex_row : ex_col;
        const double size = part == 0 ? (double)img_w : (double)img_h;
        for (int t = tid; t < ncls * d.nl; t += UFLD_THREADS) {
            const int k = t / d.nl, i = t % d.nl;
            const int stride = ncls * d.nl;
            float best = loc[t];
            int m = 0;
            for (int g = 1; g < ng; ++g) {
                const float v = loc[(size_t)g * stride + t];
                if (v > best) { best = v; m = g; }
            }
            const int valid = ex[stride + t] > ex[t] ? 1 : 0;
            const int lo = m - 1 < 0 ? 0 : m - 1;
            const int hi = m + 1 > ng - 1 ? ng - 1 : m + 1;
            // float32 softmax (x - max, exp, / sum), then float64 expectation
            float e[3];
            float sum = 0.f;
            for (int g = lo; g <= hi; ++g) {
                e[g - lo] = expf(__fsub_rn(loc[(size_t)g * stride + t], best));
                sum = __fadd_rn(sum, e[g - lo]);
            }
            double acc = 0.0;
            for (int g = lo; g <= hi; ++g) {
                const float pr = __fdiv_rn(e[g - lo], sum);
                acc = __dadd_rn(acc, __dmul_rn((double)pr, (double)g));
            }
            double c = __dadd_rn(acc, 0.5);
            c = __dmul_rn(__ddiv_rn(c, (double)(ng - 1)), size);
            s_valid[part][i][k] = valid;
            s_coord[part][i][k] = c;
        }
    }
    __syncthreads();
    if (tid < 4) {
        // output lane `tid`: 0 = col lane 0, 1 = row lane 1, 2 = row lane 2, 3 = col lane 3
        const int part = (tid == 1 || tid == 2) ? 0 : 1;
        const int lane = tid;   // head lane index equals the output slot for {0,1,2,3}
        const int ncls = part == 0 ? d.ncr : d.ncc;
        int cnt = 0;
        for (int k = 0; k < ncls; ++k) cnt += s_valid[part][lane][k];
        const bool keep = part == 0 ? ((double)cnt > (double)d.ncr / 2.0) : ((double)cnt > (double)d.ncc / 4.0);
        int n = 0;
        int32_t* o = pts + ((size_t)b * 4 + tid) * max_pts * 2;
        double* oc = coords ? coords + ((size_t)b * 4 + tid) * max_pts : nullptr;
        if (keep) {
            for (int k = 0; k < ncls; ++k) {
                if (!s_valid[part][lane][k]) continue;
                const double c = s_coord[part][lane][k];
                int px, py;
                if (part == 0) { px = (int)c; py = (int)__dmul_rn(row_anchor[k], (double)img_h); }
                else           { px = (int)__dmul_rn(col_anchor[k], (double)img_w); py = (int)c; }
                o[n * 2] = px; o[n * 2 + 1] = py;
                if (oc) oc[n] = c;
                ++n;
            }
        }
        npts[b * 4 + tid] = n;
        status[b * 4 + tid] = n > 2 ? 1 : 0;
    }
}

int launch_ufld_post(const float* heads, int ld, int B, UfldDims d, int img_w, int img_h, const double* row_anchor,
                     const double* col_anchor, int32_t* pts, int32_t* npts, uint8_t* status, double* coords, int max_pts,
                     cudaStream_t st) {
    ADAS_CHECK(d.nl == 4, "ufld_post: num_lanes must be 4 (got %d)", d.nl);
    ADAS_CHECK(d.ncr <= UFLD_MAX_ANCH && d.ncc <= UFLD_MAX_ANCH, "ufld_post: too many anchors");
    ufld_post_kernel<<<B, UFLD_THREADS, 0, st>>>(heads, ld, d, img_w, img_h, row_anchor, col_anchor, pts, npts, status, coords,
                                                max_pts);
    count_launch();
    ADAS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace adas
