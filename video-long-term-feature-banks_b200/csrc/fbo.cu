// Feature-bank kernels that work on the RAW long-term bank (HBM-bound scans / gathers), sm_100a.
//
// 1. fbo_bank_scan / fbo_bank_combine: the inference-mode Feature-Bank Operator.  Without dropout between
//    'lfb_1x1' and the phi/g projections (lfb_helper.py:320-338 with test_mode), one FBO-NL layer with a single
//    query per RoI (lfb_helper.NLCore :170-263, num_feat1 == 1) folds algebraically:
//        score_j  = theta . (W_phi (W_1 b_j + c_1) + c_phi) = (W_1^T W_phi^T theta) . b_j + const   (const drops
//                   out of the softmax),
//        sum_j p_j g_j = W_g (W_1 (sum_j p_j b_j) + c_1) + c_g                                      (sum_j p_j = 1),
//    so the layer is ONE pass over the raw bank rows b_j (R x L x D fp32): q . b_j, online softmax, weighted row
//    sum.  The (R*L x D x d) projections of the as-written graph (314.6 MMAC per RoI at L=300 + 157 MMAC per
//    layer) disappear; what is left are four R-row matmuls per layer on the tensor-core GEMM.
//    Roofline: HBM.  Algorithmic bytes per launch = R*L*D*4 (every bank element is read exactly once).
// 2. lfb_gather: builds the per-sample (L x D) bank windows on the device from a resident bank tensor and a host
//    computed row-index table (tools/lfb_loader.py:51-152 + lib/datasets/ava.py:300-323: -1 = zero padding).
#include <float.h>

#include "common.cuh"

namespace vlfb {
namespace {

constexpr int SCAN_TPB = 256;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {   // read-once data: do not keep it in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// One CTA = rows [row_begin, row_end) of one RoI.  V4 = float4 per thread per row (D = 1024 * V4), ROWS = rows per
// tile: ROWS*V4 independent 16-byte loads per thread are in flight before the first use (ROWS*D*4 = 64 KB per
// CTA, two CTAs per SM).  Per tile: partial dots -> warp shuffles -> one shared-memory exchange between the 8
// warps (double buffered: one __syncthreads per tile) -> every thread redoes the tiny online-softmax update and
// rescales / accumulates its 4*V4 columns from the registers that still hold the rows.
template <int V4, int ROWS>
__global__ void __launch_bounds__(SCAN_TPB, 2)
fbo_bank_scan_k(const float* __restrict__ bank, const float* __restrict__ q, float scale, float* __restrict__ part_acc,
                float* __restrict__ part_ml, float* __restrict__ scores, int L, int S, int rows_per_split) {
  constexpr int D = 1024 * V4;
  constexpr int NW = SCAN_TPB / 32;
  __shared__ float xch[2][NW][ROWS];
  const int r = blockIdx.x / S, sp = blockIdx.x % S;
  const int row_begin = sp * rows_per_split;
  const int row_end = min(L, row_begin + rows_per_split);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* brow = reinterpret_cast<const float4*>(bank + (int64_t)r * L * D);
  const float4* q4 = reinterpret_cast<const float4*>(q + (int64_t)r * D);
  float4 qv[V4], acc[V4];
#pragma unroll
  for (int v = 0; v < V4; ++v) {
    qv[v] = q4[v * SCAN_TPB + tid];
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float m = -FLT_MAX, l = 0.f;
  int buf = 0;
  for (int j0 = row_begin; j0 < row_end; j0 += ROWS, buf ^= 1) {
    float4 x[ROWS][V4];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const bool ok = j0 + i < row_end;
#pragma unroll
      for (int v = 0; v < V4; ++v)
        x[i][v] = ok ? ldg_stream(brow + (int64_t)(j0 + i) * (D / 4) + v * SCAN_TPB + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float part[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < V4; ++v)
        s += x[i][v].x * qv[v].x + x[i][v].y * qv[v].y + x[i][v].z * qv[v].z + x[i][v].w * qv[v].w;
      part[i] = warp_sum_f(s);
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) xch[buf][warp][i] = part[i];
    }
    __syncthreads();
    float sc[ROWS], mx = m;
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += xch[buf][w][i];
      s = (j0 + i < row_end) ? s * scale : -FLT_MAX;
      sc[i] = s;
      mx = fmaxf(mx, s);
    }
    if (scores != nullptr && tid < ROWS && j0 + tid < row_end) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += xch[buf][w][tid];
      scores[(int64_t)r * L + j0 + tid] = s * scale;
    }
    const float corr = __expf(m - mx);      // m == -FLT_MAX on the first tile: exp(-huge) == 0 and acc, l are 0
    l *= corr;
#pragma unroll
    for (int v = 0; v < V4; ++v) {
      acc[v].x *= corr; acc[v].y *= corr; acc[v].z *= corr; acc[v].w *= corr;
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const float p = (j0 + i < row_end) ? __expf(sc[i] - mx) : 0.f;
      l += p;
#pragma unroll
      for (int v = 0; v < V4; ++v) {
        acc[v].x += p * x[i][v].x; acc[v].y += p * x[i][v].y; acc[v].z += p * x[i][v].z; acc[v].w += p * x[i][v].w;
      }
    }
    m = mx;
  }
  float4* pa = reinterpret_cast<float4*>(part_acc + ((int64_t)r * S + sp) * D);
#pragma unroll
  for (int v = 0; v < V4; ++v) pa[v * SCAN_TPB + tid] = acc[v];
  if (tid == 0) {
    part_ml[((int64_t)r * S + sp) * 2 + 0] = m;
    part_ml[((int64_t)r * S + sp) * 2 + 1] = l;
  }
}

// s[r][:] = sum_sp w_sp acc_sp[:] / sum_sp w_sp l_sp, w_sp = exp(m_sp - max m); optional TF32 rounding (the result is
// the A operand of the W_1 matmul).  prob[r][j] = exp(score - M) / lsum when the scores were kept.
__global__ void fbo_bank_combine_k(const float* __restrict__ part_acc, const float* __restrict__ part_ml,
                                   float* __restrict__ out, float* __restrict__ scores_prob, int S, int D, int L,
                                   int tf32_out) {
  const int r = blockIdx.y;
  float M = -FLT_MAX;
  for (int s = 0; s < S; ++s)
    if (part_ml[((int64_t)r * S + s) * 2 + 1] > 0.f) M = fmaxf(M, part_ml[((int64_t)r * S + s) * 2]);
  float lsum = 0.f;
  for (int s = 0; s < S; ++s) {
    const float ls = part_ml[((int64_t)r * S + s) * 2 + 1];
    if (ls > 0.f) lsum += ls * __expf(part_ml[((int64_t)r * S + s) * 2] - M);
  }
  const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < D) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) {
      const float ls = part_ml[((int64_t)r * S + s) * 2 + 1];
      if (ls > 0.f) a += part_acc[((int64_t)r * S + s) * D + c] * __expf(part_ml[((int64_t)r * S + s) * 2] - M);
    }
    a *= inv;
    out[(int64_t)r * D + c] = tf32_out ? round_tf32(a) : a;
  }
  if (scores_prob != nullptr) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x)
      scores_prob[(int64_t)r * L + j] = __expf(scores_prob[(int64_t)r * L + j] - M) * inv;
  }
}

// out[i][:] = idx[i] >= 0 ? bank[idx[i]][:] : 0   (rows of D floats, D % 4 == 0); one warp-wide float4 lane per 16 B.
__global__ void lfb_gather_k(const float4* __restrict__ bank, const int32_t* __restrict__ idx, float4* __restrict__ out,
                             int64_t rows, int d4, int64_t bank_rows, int tf32_out) {
  const int64_t total = rows * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d4;
    const int c = (int)(i - row * d4);
    const int32_t src = idx[row];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src >= 0 && src < bank_rows) v = bank[(int64_t)src * d4 + c];
    if (tf32_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    out[i] = v;
  }
}

}  // namespace
}  // namespace vlfb

using namespace vlfb;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

/* Split of the L bank rows of every RoI over `S` CTAs: minimise waves x (tiles per CTA + partial-result cost) with
 * 2 resident CTAs per SM (296 per wave on B200). */
int vlfb_fbo_bank_scan_splits(int R, int L, int D) {
  if (R <= 0 || L <= 0 || (D != 1024 && D != 2048 && D != 4096)) return 0;
  const int rows = D == 4096 ? 4 : 8;
  const int max_s = (L + rows - 1) / rows;
  const int wave = 296;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= max_s && s <= 1024; ++s) {
    const int per = (L + s - 1) / s;
    if ((int64_t)(s - 1) * per >= L) continue;                 /* an empty trailing split */
    const int tiles = (per + rows - 1) / rows;
    const double waves = (double)(((int64_t)R * s + wave - 1) / wave);
    const double cost = waves * (tiles + 0.5);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

size_t vlfb_fbo_bank_scan_workspace(int R, int L, int D) {
  const int S = vlfb_fbo_bank_scan_splits(R, L, D);
  if (S <= 0) return 0;
  return ((size_t)R * S * D + (size_t)R * S * 2) * sizeof(float);
}

int vlfb_fbo_bank_scan(const float* bank, const float* q, float scale, float* out, float* prob, int R, int L, int D,
                       int tf32_out, void* workspace, size_t workspace_bytes, void* stream) {
  VLFB_CHECK_ARG(bank && q && out && R >= 0 && L > 0);
  VLFB_CHECK_ARG(D == 1024 || D == 2048 || D == 4096);
  if (R == 0) return VLFB_OK;
  const int S = vlfb_fbo_bank_scan_splits(R, L, D);
  VLFB_CHECK_ARG(S > 0 && (int64_t)R * S < (1ll << 31));
  if (workspace == nullptr || workspace_bytes < vlfb_fbo_bank_scan_workspace(R, L, D)) {
    set_error("vlfb_fbo_bank_scan: workspace of %zu bytes needed, %zu given", vlfb_fbo_bank_scan_workspace(R, L, D),
              workspace_bytes);
    return VLFB_E_WORKSPACE;
  }
  float* part_acc = static_cast<float*>(workspace);
  float* part_ml = part_acc + (size_t)R * S * D;
  const int per = (L + S - 1) / S;
  const dim3 grid((unsigned)(R * S));
  if (D == 1024)
    launch_k(fbo_bank_scan_k<1, 8>, grid, SCAN_TPB, 0, ST(stream), bank, q, scale, part_acc, part_ml, prob, L, S, per);
  else if (D == 2048)
    launch_k(fbo_bank_scan_k<2, 8>, grid, SCAN_TPB, 0, ST(stream), bank, q, scale, part_acc, part_ml, prob, L, S, per);
  else
    launch_k(fbo_bank_scan_k<4, 4>, grid, SCAN_TPB, 0, ST(stream), bank, q, scale, part_acc, part_ml, prob, L, S, per);
  VLFB_CHECK_LAUNCH();
  launch_k(fbo_bank_combine_k, dim3((unsigned)(D / 256), (unsigned)R), 256, 0, ST(stream), (const float*)part_acc,
           (const float*)part_ml, out, prob, S, D, L, tf32_out);
  VLFB_CHECK_LAUNCH();
  return VLFB_OK;
}

int vlfb_lfb_gather(const float* bank, int64_t bank_rows, const int32_t* idx, float* out, int64_t rows, int D,
                    int tf32_out, void* stream) {
  VLFB_CHECK_ARG(rows >= 0 && bank_rows >= 0 && D > 0 && (D & 3) == 0);
  if (rows == 0) return VLFB_OK;            /* an empty table: nothing to write (pointers may be NULL) */
  VLFB_CHECK_ARG(idx && out && (bank || bank_rows == 0));
  launch_k(lfb_gather_k, stream_grid(rows * (D >> 2), 256, 4), 256, 0, ST(stream), (const float4*)bank, idx, (float4*)out,
           rows, D >> 2, bank_rows, tf32_out);
  VLFB_CHECK_LAUNCH();
  return VLFB_OK;
}

}  // extern "C"
