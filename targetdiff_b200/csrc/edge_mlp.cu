// edge_mlp.cu -- the per-edge 2-layer MLPs (hk/hv/xk/xv) of the attention layers, FP32 FFMA path.
//
// Reference: BaseX2HAttLayer.forward / BaseH2XAttLayer.forward build kv_input = [edge_type | r_feat | h[dst] | h[src]]
// ([E,340], models/uni_transformer.py:45-51,111-117) and run MLP = Linear(340,128) -> LayerNorm -> ReLU -> Linear(128,out)
// (models/common.py:60-80) on it; r_feat = outer_product(edge_type, GaussianSmearing(dist)) (:194-195).
//
// Here the [E,340] tensor never exists.  With the exact first-layer split
//   pre[e] = P[dst, offA:offA+128] + P[src, offB:offB+128] + tab[type][20] + sum_j g_j(dist_e) * tab[type][j]
// where P is the node projection (node_ops.cu) and tab the gaussian/type block of the first Linear (+ bias).
// A persistent CTA (one per SM) keeps W2^T and `tab` in shared memory and loops over tiles of 128 edge slots:
//   phase 1  one warp per edge row: gather the two projected rows (coalesced 128 B segments), gaussians by
//            lanes 0..19 + shuffles, LayerNorm by shuffles, ReLU -> activation tile in shared memory
//   phase 2  128 x NOUT x 128 register-tiled FFMA GEMM out of shared memory, bias, coalesced float4 stores.
// `row_nodes` (optional) restricts the rows to the slots of a node subset (h2x only needs ligand destinations,
// because delta_x is masked to ligand atoms, models/uni_transformer.py:205-206); output rows are then compact.
#include "tdiff_common.cuh"

template <int NOUT>
__global__ void __launch_bounds__(TD_GEMM_THREADS, 1)
edge_mlp_kernel(const float* __restrict__ P, const float4* __restrict__ xm, const int* __restrict__ src,
                const unsigned char* __restrict__ etype, const int* __restrict__ row_nodes, long long n_rows, int k, TdMlp m,
                const float* __restrict__ offsets, float coeff, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                                  // [128][TD_LDA]
  float* Bs = As + 128 * TD_LDA;                     // [128][NOUT]
  float* s_tab = Bs + 128 * NOUT;                    // [4][21][128]
  float* s_g = s_tab + 4 * TD_TAB * TD_H;            // [128]
  float* s_b = s_g + TD_H;                           // [128]
  float* s_b2 = s_b + TD_H;                          // [NOUT]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 128 * NOUT / 4; i += TD_GEMM_THREADS)
    *reinterpret_cast<float4*>(Bs + 4 * i) = *reinterpret_cast<const float4*>(m.w2t + 4 * i);
  for (int i = tid; i < 4 * TD_TAB * TD_H / 4; i += TD_GEMM_THREADS)
    *reinterpret_cast<float4*>(s_tab + 4 * i) = *reinterpret_cast<const float4*>(m.tab + 4 * i);
  for (int i = tid; i < TD_H; i += TD_GEMM_THREADS) { s_g[i] = m.ln_g[i]; s_b[i] = m.ln_b[i]; }
  for (int i = tid; i < NOUT; i += TD_GEMM_THREADS) s_b2[i] = m.b2[i];
  const float mu = offsets[lane < TD_NG ? lane : 0];
  __syncthreads();

  const long long n_tiles = (n_rows + 127) / 128;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * 128;
    // ---------------- phase 1: activation tile ----------------
#pragma unroll 2
    for (int r = warp; r < 128; r += TD_GEMM_THREADS / 32) {
      const long long idx = row0 + r;
      float p[4] = {0.f, 0.f, 0.f, 0.f};
      int s = -1, dst = 0;
      long long e = 0;
      if (idx < n_rows) {
        const long long a = idx / k;
        const int j = (int)(idx - a * k);
        dst = row_nodes ? row_nodes[a] : (int)a;
        e = (long long)dst * k + j;
        s = src[e];
      }
      if (s >= 0) {
        const float4 xd = xm[dst], xs = xm[s];
        const int t = etype[e];
        const float* pa = P + (size_t)dst * TD_NPROJ + m.offA;
        const float* pb = P + (size_t)s * TD_NPROJ + m.offB;
        const float* tb = s_tab + t * (TD_TAB * TD_H);
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = (pa[lane + 32 * c] + pb[lane + 32 * c]) + tb[TD_NG * TD_H + lane + 32 * c];
        const float dx = xd.x - xs.x, dy = xd.y - xs.y, dz = xd.z - xs.z;
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        const float tt = dist - mu;
        const float gj = expf(coeff * (tt * tt));
#pragma unroll
        for (int j = 0; j < TD_NG; ++j) {
          const float g = __shfl_sync(0xffffffffu, gj, j);
#pragma unroll
          for (int c = 0; c < 4; ++c) p[c] = fmaf(g, tb[j * TD_H + lane + 32 * c], p[c]);
        }
        ln_relu_128(p, s_g, s_b, lane);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) As[r * TD_LDA + lane + 32 * c] = p[c];
    }
    __syncthreads();
    // ---------------- phase 2: second Linear ----------------
    if (NOUT == 128) {
      const int ty = tid >> 4, tx = tid & 15;
      float acc[4][8];
      tile_gemm_128(As, Bs, acc, ty, tx);
      const float4 bia0 = *reinterpret_cast<const float4*>(s_b2 + 4 * tx);
      const float4 bia1 = *reinterpret_cast<const float4*>(s_b2 + 64 + 4 * tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long idx = row0 + ty + 32 * i;
        if (idx < n_rows) {
          float* o = out + (size_t)idx * 128;
          *reinterpret_cast<float4*>(o + 4 * tx) = make_float4(acc[i][0] + bia0.x, acc[i][1] + bia0.y, acc[i][2] + bia0.z, acc[i][3] + bia0.w);
          *reinterpret_cast<float4*>(o + 64 + 4 * tx) = make_float4(acc[i][4] + bia1.x, acc[i][5] + bia1.y, acc[i][6] + bia1.z, acc[i][7] + bia1.w);
        }
      }
    } else {   // NOUT == 16: 128 x 16 outputs, 4 per thread
      const int r = tid >> 2, c4 = tid & 3;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int kk = 0; kk < 128; kk += 4) {
        const float4 a = *reinterpret_cast<const float4*>(As + r * TD_LDA + kk);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(Bs + (kk + q) * NOUT + 4 * c4);
          acc.x = fmaf(av[q], b.x, acc.x); acc.y = fmaf(av[q], b.y, acc.y);
          acc.z = fmaf(av[q], b.z, acc.z); acc.w = fmaf(av[q], b.w, acc.w);
        }
      }
      const long long idx = row0 + r;
      if (idx < n_rows) {
        const float4 bia = *reinterpret_cast<const float4*>(s_b2 + 4 * c4);
        *reinterpret_cast<float4*>(out + (size_t)idx * NOUT + 4 * c4) = make_float4(acc.x + bia.x, acc.y + bia.y, acc.z + bia.z, acc.w + bia.w);
      }
    }
    __syncthreads();
  }
}

template <int NOUT>
static void launch_impl(const float* P, const float4* xm, const int* src, const unsigned char* etype, const int* row_nodes,
                        long long n_rows, int k, TdMlp m, const float* offsets, float coeff, float* out, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)(128 * TD_LDA + 128 * NOUT + 4 * TD_TAB * TD_H + 2 * TD_H + NOUT) * sizeof(float);
  static size_t opted[TD_MAX_DEVICES] = {0};
  td_opt_in_smem(edge_mlp_kernel<NOUT>, smem, opted);
  long long n_tiles = (n_rows + 127) / 128;
  int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
  edge_mlp_kernel<NOUT><<<grid, TD_GEMM_THREADS, smem, st>>>(P, xm, src, etype, row_nodes, n_rows, k, m, offsets, coeff, out);
}

void td_launch_edge_mlp(const float* P, const float4* xm, const int* src, const unsigned char* etype, const int* row_nodes,
                        long long n_rows, int k, TdMlp m, const float* offsets, float coeff, float* out, int sm_count, cudaStream_t st) {
  if (n_rows == 0) return;
  if (m.nout == 128) launch_impl<128>(P, xm, src, etype, row_nodes, n_rows, k, m, offsets, coeff, out, sm_count, st);
  else launch_impl<16>(P, xm, src, etype, row_nodes, n_rows, k, m, offsets, coeff, out, sm_count, st);
}
