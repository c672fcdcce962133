// knn.cu -- fused per-graph k-NN -> dst-sorted fixed-degree neighbour list (CSR with constant row length k).
//
// Replaces torch_geometric.nn.knn_graph(x, k, batch, flow='source_to_target') -> torch_cluster.knn
// (reference call site models/uni_transformer.py:280).  Canonical semantics:
//   d2 = ((dx*dx)+(dy*dy))+(dz*dz), every op rounded to fp32 (no FMA contraction); ascending d2, ties -> smaller
//   node index; self removed; graphs with <= k nodes give fewer edges (slots padded with -1).
// The (d2, index) order is realised with one 64-bit key  (float_as_uint(d2) << 32) | local_index : d2 >= 0 so the
// IEEE bit pattern is monotone, keys are unique, and "k+1 smallest keys" is exactly the canonical selection.
//
// Mapping: one CTA per (graph, chunk of queries); the graph's coordinates are staged once in shared memory
// (coalesced float4 loads); one warp per query holds the key row in shared memory, each lane tracks the minimum of
// its strided slice; k+1 rounds of a 64-bit warp-min pick the neighbours in order, only the winning lane rescans.
#include "tdiff_common.cuh"

#define KNN_WARPS 8

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w < v) ? w : v;
  }
  return v;
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(const float4* __restrict__ xm, const int* __restrict__ node_ptr, int k, int max_ng, int* __restrict__ src) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* spos = reinterpret_cast<float4*>(smem_raw);                                     // [max_ng]
  unsigned long long* skeys = reinterpret_cast<unsigned long long*>(spos + max_ng);       // [KNN_WARPS][max_ng]

  const int g = blockIdx.x;
  const int base = node_ptr[g];
  const int ng = node_ptr[g + 1] - base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < ng; j += blockDim.x) spos[j] = xm[base + j];
  __syncthreads();

  unsigned long long* keys = skeys + (size_t)warp * max_ng;
  const int rounds = min(k + 1, ng);
  // queries of this graph are spread over gridDim.y chunks and the CTA's warps
  for (int qi = blockIdx.y * KNN_WARPS + warp; qi < ng; qi += gridDim.y * KNN_WARPS) {
    const float4 xq = spos[qi];
    unsigned long long lmin = ~0ull;
    for (int j = lane; j < ng; j += 32) {
      const float4 xc = spos[j];
      const float dx = __fsub_rn(xq.x, xc.x), dy = __fsub_rn(xq.y, xc.y), dz = __fsub_rn(xq.z, xc.z);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
      keys[j] = key;
      lmin = key < lmin ? key : lmin;
    }
    __syncwarp();
    int* out = src + (size_t)(base + qi) * k;
    int written = 0;
    for (int r = 0; r < rounds; ++r) {
      const unsigned long long gmin = warp_min_u64(lmin);
      const int j = (int)(gmin & 0xffffffffu);
      if (lmin == gmin) {               // unique winner lane: retire the key and rescan its slice
        keys[j] = ~0ull;
        unsigned long long m = ~0ull;
        for (int jj = lane; jj < ng; jj += 32) {
          const unsigned long long kv = keys[jj];
          m = kv < m ? kv : m;
        }
        lmin = m;
      }
      if (j != qi && written < k) {
        if (lane == 0) out[written] = base + j;
        ++written;
      }
    }
    for (int w = written + lane; w < k; w += 32) out[w] = -1;
    __syncwarp();
  }
}

void td_launch_knn(const float4* xm, const int* node_ptr, int n_graphs, int max_ng, int k, int* src, cudaStream_t st) {
  if (n_graphs <= 0) return;
  size_t smem = (size_t)max_ng * sizeof(float4) + (size_t)KNN_WARPS * max_ng * sizeof(unsigned long long);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  // enough chunks that small batches still fill the 148 SMs; each chunk re-stages the coordinates (cheap)
  int chunks = (max_ng + KNN_WARPS * 8 - 1) / (KNN_WARPS * 8);
  if (chunks < 1) chunks = 1;
  if ((long long)n_graphs * chunks > 65535LL * 8) chunks = 1;
  dim3 grid(n_graphs, chunks);
  knn_kernel<<<grid, KNN_WARPS * 32, smem, st>>>(xm, node_ptr, k, max_ng, src);
}
