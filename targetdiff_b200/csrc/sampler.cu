// sampler.cu -- the per-step diffusion update fused into one kernel, plus small state/marshalling kernels.
//
// step_epilogue restates the loop body after the network call of ScorePosNet3D.sample_diffusion
// (reference models/molopt_score_model.py:667-693, C0 mode) and its helpers:
//   q_pos_posterior :424-428, extract :706-708, index_to_log_onehot :124-130 (clamp 1e-30), log_add_exp :173-175,
//   q_v_pred :383-392, q_v_pred_one_timestep :371-381, q_v_posterior :401-409, log_sample_categorical :160-166.
// The reference issues ~25 elementwise launches and 4 D2H copies per step; here it is one launch, trajectories are
// written straight into preallocated device buffers, and the step index lives in device memory so that the whole
// step can be replayed from a CUDA graph with zero host synchronisation.
#include "tdiff_common.cuh"
#include "sampler.cuh"

// ---------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(unsigned x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0,1), 24 bits

__device__ __forceinline__ float log_add_exp_f(float a, float b) {
  const float mx = fmaxf(a, b);
  return mx + logf(expf(a - mx) + expf(b - mx));
}

__global__ void step_epilogue_kernel(TdStepArgs A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.n_lig) return;
  const int s = *A.step;                       // steps done so far
  const int t = A.t_start - s;                 // current timestep (reference :649-651)
  const int K = A.n_classes;

  // noise for this atom
  float nz[3];
  float un[TD_CMAX];
  if (A.pos_noise) {
    const float* pn = A.pos_noise + ((size_t)s * A.n_lig + a) * 3;
    nz[0] = pn[0]; nz[1] = pn[1]; nz[2] = pn[2];
  } else {
    const uint2 key = make_uint2((unsigned)A.seed, (unsigned)(A.seed >> 32));
    const uint4 r0 = philox4x32_10(make_uint4((unsigned)a, (unsigned)s, 0u, 0x70737400u), key);
    // Box-Muller on (0,1] uniforms
    const float u0 = 1.0f - u01(r0.x), u1 = u01(r0.y), u2 = 1.0f - u01(r0.z), u3 = u01(r0.w);
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    nz[0] = ra * cospif(2.0f * u1); nz[1] = ra * sinpif(2.0f * u1); nz[2] = rb * cospif(2.0f * u3);
  }
  if (!A.pos_only) {
    if (A.v_uniform) {
      const float* vu = A.v_uniform + ((size_t)s * A.n_lig + a) * K;
      for (int c = 0; c < K; ++c) un[c] = vu[c];
    } else {
      const uint2 key = make_uint2((unsigned)A.seed, (unsigned)(A.seed >> 32));
      for (int c0 = 0; c0 < K; c0 += 4) {
        const uint4 r = philox4x32_10(make_uint4((unsigned)a, (unsigned)s, 1u + (unsigned)(c0 >> 2), 0x76756e69u), key);
        un[c0] = u01(r.x);
        if (c0 + 1 < K) un[c0 + 1] = u01(r.y);
        if (c0 + 2 < K) un[c0 + 2] = u01(r.z);
        if (c0 + 3 < K) un[c0 + 3] = u01(r.w);
      }
    }
  }

  // ---- positions: posterior mean + noise (reference :673-679)
  const int g = A.lig_graph[a];
  const float4 xt = A.lig_pos[a];
  float4 x0 = A.xm_final[A.lig_node[a]];
  if (A.mean_noise) {              // x0 = sqrt(1/ac) x_t - sqrt(1/ac - 1) (pred - x_t)   (reference :419-422,663-666)
    const float ra = A.sra[t], rm = A.srm1[t];
    x0.x = ra * xt.x - rm * (x0.x - xt.x);
    x0.y = ra * xt.y - rm * (x0.y - xt.y);
    x0.z = ra * xt.z - rm * (x0.z - xt.z);
  }
  const float c0 = A.c0[t], ct = A.ct[t];
  const float sig = ((t == 0) ? 0.0f : 1.0f) * expf(0.5f * A.logvar[t]);
  float4 xn;
  xn.x = (c0 * x0.x + ct * xt.x) + sig * nz[0];
  xn.y = (c0 * x0.y + ct * xt.y) + sig * nz[1];
  xn.z = (c0 * x0.z + ct * xt.z) + sig * nz[2];
  xn.w = 1.0f;
  A.lig_pos[a] = xn;
  const float4 off = A.offset[g];
  if (A.pos_traj) {
    float* o = A.pos_traj + ((size_t)s * A.n_lig + a) * 3;
    o[0] = xn.x + off.x; o[1] = xn.y + off.y; o[2] = xn.z + off.z;
  }

  int vnew = A.lig_v[a];
  if (!A.pos_only) {
    // ---- atom types: categorical posterior in log space (reference :682-685)
    const float* lg = A.logits + (size_t)a * K;
    float lr[TD_CMAX];
    float mx = -INFINITY;
    for (int c = 0; c < K; ++c) { lr[c] = lg[c]; mx = fmaxf(mx, lr[c]); }
    float se = 0.0f;
    for (int c = 0; c < K; ++c) se += expf(lr[c] - mx);
    const float lse = logf(se);
    for (int c = 0; c < K; ++c) lr[c] = (lr[c] - mx) - lse;                 // log_softmax
    const int tm1 = t > 0 ? t - 1 : 0;
    const float lca = A.lca_v[tm1], l1mca = A.l1mca_v[tm1] - A.log_k;
    const float la = A.la_v[t], l1ma = A.l1ma_v[t] - A.log_k;
    const int vcur = vnew;
    const float log_eps = -69.07755279f;                                   // logf(1e-30f)
    float un_lp[TD_CMAX];
    float m2 = -INFINITY;
    for (int c = 0; c < K; ++c) {
      const float lvt = (c == vcur) ? 0.0f : log_eps;
      un_lp[c] = log_add_exp_f(lr[c] + lca, l1mca) + log_add_exp_f(lvt + la, l1ma);
      m2 = fmaxf(m2, un_lp[c]);
    }
    float s2 = 0.0f;
    for (int c = 0; c < K; ++c) s2 += expf(un_lp[c] - m2);
    const float lse2 = m2 + logf(s2);                                      // torch.logsumexp
    float best = -INFINITY;
    vnew = 0;
    float* o0 = A.v0_traj ? A.v0_traj + ((size_t)s * A.n_lig + a) * K : nullptr;
    float* ot = A.vt_traj ? A.vt_traj + ((size_t)s * A.n_lig + a) * K : nullptr;
    for (int c = 0; c < K; ++c) {
      const float lp = un_lp[c] - lse2;
      const float gum = -logf(-logf(un[c] + 1e-30f) + 1e-30f);
      const float sc = gum + lp;
      if (sc > best) { best = sc; vnew = c; }
      if (o0) o0[c] = lr[c];
      if (ot) ot[c] = lp;
    }
    A.lig_v[a] = vnew;
  }
  if (A.v_traj) A.v_traj[(size_t)s * A.n_lig + a] = (long long)vnew;
}

__global__ void advance_step_kernel(int* step) { *step += 1; }

void td_launch_step_epilogue(const TdStepArgs& A, cudaStream_t st) {
  if (A.n_lig > 0) step_epilogue_kernel<<<(A.n_lig + 127) / 128, 128, 0, st>>>(A);
  advance_step_kernel<<<1, 1, 0, st>>>(A.step);
}

// ---------------------------------------------------------------------------------------- state marshalling
// per-graph protein centroid, sequential fp32 sum in atom order then / count == torch_scatter.scatter_mean on CPU
// (reference models/molopt_score_model.py:115)
__global__ void segment_mean3_kernel(const float* __restrict__ pos, const int* __restrict__ seg_ptr, int n_seg, float4* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_seg) return;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  const int b = seg_ptr[g], e = seg_ptr[g + 1];
  for (int i = b; i < e; ++i) { sx += pos[3 * i]; sy += pos[3 * i + 1]; sz += pos[3 * i + 2]; }
  float cnt = (float)(e - b);
  if (cnt < 1.0f) cnt = 1.0f;
  out[g] = make_float4(sx / cnt, sy / cnt, sz / cnt, 0.0f);
}
void td_launch_segment_mean3(const float* pos, const int* seg_ptr, int n_seg, float4* out, cudaStream_t st) {
  if (n_seg > 0) segment_mean3_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(pos, seg_ptr, n_seg, out);
}

// protein atoms -> node array (both ping-pong buffers), centred
__global__ void place_protein_kernel(const float* __restrict__ pos, const int* __restrict__ prot_node, const int* __restrict__ prot_graph,
                                     const float4* __restrict__ offset, int n, float4* __restrict__ xm0, float4* __restrict__ xm1) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float4 o = offset[prot_graph[p]];
  const float4 v = make_float4(pos[3 * p] - o.x, pos[3 * p + 1] - o.y, pos[3 * p + 2] - o.z, 0.0f);
  xm0[prot_node[p]] = v;
  xm1[prot_node[p]] = v;
}
void td_launch_place_protein(const float* pos, const int* prot_node, const int* prot_graph, const float4* offset, int n, float4* xm0,
                             float4* xm1, cudaStream_t st) {
  if (n > 0) place_protein_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, prot_node, prot_graph, offset, n, xm0, xm1);
}

// ligand state in  (lab frame -> centred float4, int64 -> int32, range check like reference :125)
__global__ void set_ligand_kernel(const float* __restrict__ pos, const long long* __restrict__ v, const int* __restrict__ lig_graph,
                                  const float4* __restrict__ offset, int apply_center, int n, int n_classes, float4* __restrict__ lig_pos,
                                  int* __restrict__ lig_v, int* __restrict__ err) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (apply_center) o = offset[lig_graph[a]];
  lig_pos[a] = make_float4(pos[3 * a] - o.x, pos[3 * a + 1] - o.y, pos[3 * a + 2] - o.z, 1.0f);
  if (v) {
    const long long vv = v[a];
    if (vv < 0 || vv >= n_classes) { atomicExch(err, 1); lig_v[a] = 0; }      // never store an index that would read out of bounds
    else lig_v[a] = (int)vv;
  }
}
void td_launch_set_ligand(const float* pos, const long long* v, const int* lig_graph, const float4* offset, int apply_center, int n,
                          int n_classes, float4* lig_pos, int* lig_v, int* err, cudaStream_t st) {
  if (n > 0) set_ligand_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, v, lig_graph, offset, apply_center, n, n_classes, lig_pos, lig_v, err);
}

__global__ void get_ligand_kernel(const float4* __restrict__ lig_pos, const int* __restrict__ lig_v, const int* __restrict__ lig_graph,
                                  const float4* __restrict__ offset, int add_offset, int n, float* __restrict__ pos, long long* __restrict__ v) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  if (pos) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (add_offset) o = offset[lig_graph[a]];
    const float4 p = lig_pos[a];
    pos[3 * a] = p.x + o.x; pos[3 * a + 1] = p.y + o.y; pos[3 * a + 2] = p.z + o.z;
  }
  if (v) v[a] = (long long)lig_v[a];
}
void td_launch_get_ligand(const float4* lig_pos, const int* lig_v, const int* lig_graph, const float4* offset, int add_offset, int n,
                          float* pos, long long* v, cudaStream_t st) {
  if (n > 0) get_ligand_kernel<<<(n + 255) / 256, 256, 0, st>>>(lig_pos, lig_v, lig_graph, offset, add_offset, n, pos, v);
}

// ligand atoms parked far from every pocket (distinct points; ligand-free cache construction, engine.cu)
__global__ void park_ligand_kernel(float4* __restrict__ lig_pos, int* __restrict__ lig_v, int n) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  lig_pos[a] = make_float4(1.0e6f + 64.0f * (float)(a & 4095), 1.0e6f + 64.0f * (float)((a >> 12) & 4095), 1.0e6f, 1.0f);
  lig_v[a] = 0;
}
void td_launch_park_ligand(float4* lig_pos, int* lig_v, int n, cudaStream_t st) {
  if (n > 0) park_ligand_kernel<<<(n + 255) / 256, 256, 0, st>>>(lig_pos, lig_v, n);
}

// ligand rows of the node array <- ligand state (start of every forward)
__global__ void scatter_ligand_pos_kernel(const float4* __restrict__ lig_pos, const int* __restrict__ lig_node, int n, float4* __restrict__ xm) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n) xm[lig_node[a]] = lig_pos[a];
}
void td_launch_scatter_ligand_pos(const float4* lig_pos, const int* lig_node, int n, float4* xm, cudaStream_t st) {
  if (n > 0) scatter_ligand_pos_kernel<<<(n + 255) / 256, 256, 0, st>>>(lig_pos, lig_node, n, xm);
}

// float4 node rows -> packed [n,3] (all nodes when idx == NULL, else gathered rows)
__global__ void gather_xyz_kernel(const float4* __restrict__ xm, const int* __restrict__ idx, int n, float* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const float4 p = xm[idx ? idx[a] : a];
  out[3 * a] = p.x; out[3 * a + 1] = p.y; out[3 * a + 2] = p.z;
}
void td_launch_gather_xyz(const float4* xm, const int* idx, int n, float* out, cudaStream_t st) {
  if (n > 0) gather_xyz_kernel<<<(n + 255) / 256, 256, 0, st>>>(xm, idx, n, out);
}

__global__ void pack_xyzm_kernel(const float* __restrict__ x, const unsigned char* __restrict__ mask, int n, float4* __restrict__ xm) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n) xm[a] = make_float4(x[3 * a], x[3 * a + 1], x[3 * a + 2], (mask && mask[a]) ? 1.0f : 0.0f);
}
void td_launch_pack_xyzm(const float* x, const unsigned char* mask, int n, float4* xm, cudaStream_t st) {
  if (n > 0) pack_xyzm_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, mask, n, xm);
}

// ---------------------------------------------------------------------------------------- edge_index export
// slots [N*k] (-1 padded) -> compact int64 [2,E] in slot order (dst ascending, then distance ascending):
// exactly PyG knn_graph(flow='source_to_target') row0 = src, row1 = dst.  Single-CTA scan over per-node degrees.
__global__ void __launch_bounds__(1024)
edge_count_scan_kernel(const int* __restrict__ src, int n_nodes, int k, long long* __restrict__ node_off, long long* __restrict__ total) {
  __shared__ long long part[1024];
  const int tid = threadIdx.x;
  const int per = (n_nodes + 1023) / 1024;
  const int b = min(tid * per, n_nodes), e = min(b + per, n_nodes);
  long long s = 0;
  for (int n = b; n < e; ++n) {
    int d = 0;
    for (int j = 0; j < k; ++j) d += (src[(size_t)n * k + j] >= 0);
    s += d;
  }
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    long long run = 0;
    for (int i = 0; i < 1024; ++i) { const long long v = part[i]; part[i] = run; run += v; }
    *total = run;
  }
  __syncthreads();
  long long run = part[tid];
  for (int n = b; n < e; ++n) {
    node_off[n] = run;
    int d = 0;
    for (int j = 0; j < k; ++j) d += (src[(size_t)n * k + j] >= 0);
    run += d;
  }
}
__global__ void edge_compact_kernel(const int* __restrict__ src, const float* __restrict__ e_w, int n_nodes, int k,
                                    const long long* __restrict__ node_off, const long long* __restrict__ total,
                                    long long* __restrict__ edge_index, float* __restrict__ ew_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_nodes * k) return;
  const int n = (int)(i / k), j = (int)(i % k);
  const int s = src[i];
  if (s < 0) return;
  const long long E = *total;
  const long long pos = node_off[n] + j;       // valid slots are leading
  if (edge_index) { edge_index[pos] = s; edge_index[E + pos] = n; }
  if (ew_out) ew_out[pos] = e_w[i];
}
void td_launch_edge_count_scan(const int* src, int n_nodes, int k, long long* node_off, long long* total, cudaStream_t st) {
  edge_count_scan_kernel<<<1, 1024, 0, st>>>(src, n_nodes, k, node_off, total);
}
void td_launch_edge_compact(const int* src, const float* e_w, int n_nodes, int k, const long long* node_off, const long long* total,
                            long long* edge_index, float* ew_out, cudaStream_t st) {
  const long long n = (long long)n_nodes * k;
  if (n > 0) edge_compact_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(src, e_w, n_nodes, k, node_off, total, edge_index, ew_out);
}
