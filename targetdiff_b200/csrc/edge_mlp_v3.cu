// edge_mlp_v3.cu -- per-edge MLP, third generation: BOTH Linear layers' dense parts on tcgen05, LayerNorm thread-per-row.
//
// Math (identical to edge_mlp.cu / edge_mlp_tc.cu; reference models/uni_transformer.py:45-56,111-120, models/common.py:60-80):
//   pre[e]  = P[dst, offA:+128] + P[src, offB:+128] + tab[type][20] + sum_j g_j(dist_e) * tab[type][j]
//   hid     = relu(LN(pre) * ln_g + ln_b)
//   out[e]  = hid . W2^T + b2
//
// Why a third version: in edge_mlp_tc.cu the CUDA-core "producers" (lane = feature) spend ~140 warp-instructions per edge
// row, half of them on the 20x128 gaussian block (FFMA2 + a shuffle per (row, gaussian)) and a quarter on shuffle-based
// LayerNorm; the tensor pipe idles at ~10 %.  Here
//   * the gaussian/type block of protein-protein edges (type 3, 93 % of all edges) is a second, small MMA:
//       Dpre[128 x 128] = G[128 x 32] . Tab3^T,   G row = (g_0..g_19, 1, 0...) split in bf16 pieces,
//     accumulated in TMEM and read back with tcgen05.ld -- thread i of a warp owns accumulator row 32q+i, which is
//     exactly the layout a thread-per-row LayerNorm wants (no shuffles: a row's 128 features live in 2 threads);
//   * the coalesced gathers stay lane = feature, but in dedicated warps that only move data: P[src] (+ P[dst], + the rare
//     non-type-3 gaussian block on CUDA cores) -> a swizzled fp32 staging tile in shared memory; row threads pick their
//     row up from there (conflict-free both ways).
//
// CTA = 28 warps, one CTA per SM, persistent over tiles of 128 edge slots:
//   warps  0-3   epilogue      TMEM D -> +b2 -> global                                   (setmaxnreg 64)
//   warp   4     MMA issuer    Dpre = G.Tab3^T ; D = A.W2^T (tcgen05.mma, one thread)     (56; warps 5-7 idle)
//   warps  8-11  gather        32 rows each: cp.async 512 B rows of P[src] -> S           (40)
//   warps 12-27  row threads   warp 12+q+4*qq: rows 32q..32q+31, feature quarter qq        (80)
//                              gaussians -> G pieces; P[dst] + S + Dpre -> LayerNorm (4-thread exchange through smem + named
//                              barriers, packed f32x2 math) -> ReLU -> bf16 split -> A pieces (UMMA K-major SWIZZLE_128B)
// Shared memory (226 KB): W2 pieces 64 KB | A pieces 64 KB | S fp32 64 KB | G pieces 16 KB | Tab3 pieces 16 KB | 2 KB misc.
// TMEM 512 columns: D[2] at 0/128, Dpre[2] at 256/384.  bf16 split: 2 pieces / 3 products (see edge_mlp_tc.cu).
#include <stdio.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#include "tdiff_common.cuh"

namespace v3 {

// Warp roles (warpgroup-aligned for setmaxnreg): 0-3 epilogue of accumulator columns 0-63, 4-7 epilogue of columns 64-127 (warp w
// reads TMEM lanes 32*(w%4) ..), 8-11 gather (11 also issues the MMAs), 12-27 row threads.
constexpr int kThreads = 28 * 32;
constexpr int kEpiWarps = 8, kGatherWarp0 = 8, kGatherWarps = 4, kMmaWarp = 11, kRowWarp0 = 12, kRowWarps = 16;
constexpr int kPiece = 128 * 128 * 2;      // bf16 piece of a 128x128 tile (two K-halves of 128 rows x 128 B)
constexpr int kAtom = 128 * 128;           // 128 rows x 128 B
constexpr int kGPiece = 128 * 64;          // bf16 piece of a 128 x 32 tile (SWIZZLE_64B, 64 B rows)
constexpr int kSBytes = 128 * 128 * 4;     // fp32 staging tile: 4 column atoms of 128 rows x 128 B
// shared-memory map (bytes from the 1024-aligned base)
constexpr int oW = 0, oA = oW + 2 * kPiece, oS = oA + 2 * kPiece, oG = oS + kSBytes, oT = oG + 2 * kGPiece, oX = oT + 2 * kGPiece,
              oBar = oX + 4 * 128 * 4, kSmem = oBar + 16 * 8 + 16;
enum { B_S_FULL = 0, B_S_EMPTY, B_G_FULL, B_A_FULL, B_A_EMPTY, B_DPRE_FULL0, B_DPRE_FULL1, B_D_FULL0, B_D_FULL1, B_D_EMPTY0, B_D_EMPTY1 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(100);
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major UMMA descriptors: SWIZZLE_128B (8-row groups 1024 B apart) and SWIZZLE_64B (8-row groups 512 B apart)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // bf16 x bf16 -> f32, M=N=128

__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// 16-byte asynchronous global -> shared copy (LDGSTS, L2-only caching), no register staging
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 32-byte global store (one full sector per thread)
__device__ __forceinline__ void stg256(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}
// packed fp32 pairs (Blackwell FADD2 / FFMA2)
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// line prefetch into L1 (the kernel's own gathers are L2-only cp.async copies, so the small L1 is left to these broadcast rows)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 32-byte streaming load (one full sector per thread, not kept in L1): row-strided per-thread reads would otherwise pull whole
// 128-byte lines through a 28 KB L1 once per 16-byte piece
__device__ __forceinline__ void ldg256_stream(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// LayerNorm affine parameters travel as a kernel argument (constant bank: no L2 round trips in the row threads)
struct LnParams { float g[128]; float b[128]; float b2[128]; };
// Fused attention aggregation in the value-MLP epilogue (k == 32: the 32 rows of an epilogue warp are exactly the edges of one
// destination): h[dst] += sum_e softmax_e(logits[e,head]) * e_w[e] * v[e]   (reference models/uni_transformer.py:73-83).
struct AggArgs {
  const float* logits;   // [E,16] written by the key-MLP launch; NULL = plain value output
  const float* e_w;      // [E]
  float* h;              // [N,128] node features, updated in place (row of a destination is touched by one warp only)
  int n_nodes;
  int key_softmax;       // key launch (k == 32): write softmax(logits) * e_w instead of the raw logits; the value launch then reads weights
};

// Reduce N (8 or 16) per-lane values over the 32 lanes of a warp with a transposing butterfly: N - 1 + log2(32 / N) shuffles instead
// of 5 N.  On return lane l holds the total (sum or max) of element (l & (N - 1)).
template <int N, bool MAX>
__device__ __forceinline__ float warp_transpose_reduce(float (&t)[N], int lane) {
  auto op = [](float a, float b) { return MAX ? fmaxf(a, b) : a + b; };
#pragma unroll
  for (int h = N / 2; h >= 1; h >>= 1) {
    const bool up = lane & h;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? t[i] : t[i + h];
      const float keep = up ? t[i + h] : t[i];
      t[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, h));
    }
  }
  float r = t[0];
#pragma unroll
  for (int m = N; m < 32; m <<= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, m));
  return r;
}
// 8 consecutive values -> two bf16 pieces (16 bytes each); residual of the first piece is exact in fp32
__device__ __forceinline__ void split8_store(uint32_t addr_p0, uint32_t addr_p1, const float (&y)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = cvt_bf16x2(y[2 * i + 1], y[2 * i]);
    float r0, r1;
    upk2(sub2(pk2(y[2 * i], y[2 * i + 1]), pk2(__uint_as_float(h[i] << 16), __uint_as_float(h[i] & 0xffff0000u))), r0, r1);
    l[i] = cvt_bf16x2(r1, r0);
  }
  sts128(addr_p0, h[0], h[1], h[2], h[3]);
  sts128(addr_p1, l[0], l[1], l[2], l[3]);
}

}  // namespace v3

using namespace v3;

// NOUT = 128: key / value MLPs (hk, hv, xk);  NOUT = 16: the per-head scalar value MLP of h2x (xv)
template <int NOUT>
__global__ void __launch_bounds__(kThreads, 1)
edge_mlp_v3_kernel(const float* __restrict__ P, const int* __restrict__ src, const unsigned char* __restrict__ etype,
                   const float* __restrict__ dist_arr, const int* __restrict__ row_nodes, long long n_rows, int k, TdMlp m,
                   const unsigned char* __restrict__ w2_image, const unsigned char* __restrict__ tab3_image, const float* __restrict__ offsets,
                   float coeff, const float* __restrict__ tslow, const float* __restrict__ qnode, float* __restrict__ out, AggArgs agg, const int* __restrict__ d_n_dst, int dbg, const __grid_constant__ LnParams lp, long long* __restrict__ ts) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sW = sbase + oW, sA = sbase + oA, sS = sbase + oS, sG = sbase + oG, sT = sbase + oT, sX = sbase + oX, sBar = sbase + oBar;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem_raw + oBar + 16 * 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };
  // debug timeline: role `ro`, event `ev` of local tile `t` (CTA 0, first 16 tiles) -> ts[(t*4 + ro)*8 + ev]
  auto stamp = [&](int ro, long long t, int ev) {
#ifdef TDIFF_V3_TIMELINE
    if (ts && blockIdx.x == 0 && t < 16 && (threadIdx.x & 31) == 0) ts[(t * 4 + ro) * 8 + ev] = clock64();
#endif
  };
  // row -> (destination slot, neighbour slot): k is a power of two for every shipped configuration but 48
  const int kshift = (k & (k - 1)) == 0 ? __ffs(k) - 1 : -1;
  auto row_dst = [&](long long idx, int& j) -> unsigned {
    const unsigned a = kshift >= 0 ? (unsigned)idx >> kshift : (unsigned)idx / (unsigned)k;
    j = (int)((unsigned)idx - a * (unsigned)k);
    return a;
  };

  if ((sbase & 1023u) != 0) __trap();            // SWIZZLE_128B atoms need a 1024-byte aligned window
  // ---- one-time setup: weight images -> smem, barriers, TMEM
  constexpr int kWAtom = NOUT * 128;            // one K-half of a weight piece: NOUT rows x 128 B
  constexpr int kWPiece = 2 * kWAtom;
  constexpr uint32_t kIdescMain = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
  for (int i = tid; i < 2 * kWPiece / 16; i += kThreads) {
    const uint4 v = reinterpret_cast<const uint4*>(w2_image)[i];
    sts128(sW + 16 * i, v.x, v.y, v.z, v.w);
  }
  for (int i = tid; i < 2 * kGPiece / 16; i += kThreads) {
    const uint4 v = reinterpret_cast<const uint4*>(tab3_image)[i];
    sts128(sT + 16 * i, v.x, v.y, v.z, v.w);
  }
  if (tid == 0) {
    mbar_init(bar(B_S_FULL), kGatherWarps);
    mbar_init(bar(B_S_EMPTY), kRowWarps);
    mbar_init(bar(B_G_FULL), kRowWarps);
    mbar_init(bar(B_A_FULL), kRowWarps);
    mbar_init(bar(B_A_EMPTY), 1);
    mbar_init(bar(B_DPRE_FULL0), 1);
    mbar_init(bar(B_DPRE_FULL1), 1);
    mbar_init(bar(B_D_FULL0), 1);
    mbar_init(bar(B_D_FULL1), 1);
    mbar_init(bar(B_D_EMPTY0), kEpiWarps);
    mbar_init(bar(B_D_EMPTY1), kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(s_tmem), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (d_n_dst) n_rows = (long long)(*d_n_dst) * k;        // destination subset compacted on the device (row_nodes list)
  const long long n_tiles = (n_rows + 127) / 128;
  const long long my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp >= kRowWarp0) {
    // ================================================================= row threads (thread = edge row x 32 features)
    reg_inc<80>();
    const int rwp = warp - kRowWarp0, q = rwp & 3, qq = rwp >> 2;
    const int r = 32 * q + lane;                    // row of the tile == TMEM lane
    float mu[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) mu[i] = offsets[5 * qq + i];
    const float coeff2 = coeff * 1.4426950408889634f;
    // metadata of this thread's row in tile `t` (s < 0: absent edge / beyond n_rows)
    auto load_md = [&](long long t, int& s_, int& ty_, int& dst_, float& dist_) {
      s_ = -1; ty_ = 0; dst_ = 0; dist_ = 0.f;
      if (t < my_tiles) {
        const long long idx = (blockIdx.x + t * (long long)gridDim.x) * 128 + r;
        if (idx < n_rows) {
          int j;
          const unsigned a = row_dst(idx, j);
          dst_ = row_nodes ? row_nodes[a] : (int)a;
          const size_t e = (size_t)dst_ * k + j;
          s_ = src[e]; ty_ = etype[e]; dist_ = dist_arr[e];
        }
      }
    };
    // gaussian chunk of G for one tile.  K slots: 8*qq + i = gaussian 5*qq + i (i < 5); slot 29 = 1 (constant row); others 0.
    // Only type-3 rows are non-zero (other types are completed by the gather warps on CUDA cores).
    auto write_g = [&](int s_, int ty_, float dist_) {
      const bool t3 = s_ >= 0 && ty_ == 3;
      float gv[8];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float t = dist_ - mu[i];
        gv[i] = t3 ? ex2_approx(coeff2 * (t * t)) : 0.0f;          // exp(coeff t^2); the bf16 split below keeps 16 bits of it
      }
      gv[5] = (t3 && qq == 3) ? 1.0f : 0.0f;
      gv[6] = gv[7] = 0.0f;
      const uint32_t a0 = sG + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u + (((uint32_t)qq ^ (uint32_t)((r >> 1) & 3)) << 4);
      split8_store(a0, a0 + kGPiece, gv);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_G_FULL));
    };
    int s0, t0, d0, s1, t1, d1;
    float dist0, dist1;
    load_md(0, s0, t0, d0, dist0);
    if (my_tiles > 0) write_g(s0, t0, dist0);
    load_md(1, s1, t1, d1, dist1);
    const uint32_t xslot = sX + (uint32_t)r * 4u;          // exchange slot of this row; quarter qq at + qq*512
    for (long long it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const bool valid = s0 >= 0;
      if (rwp == 0) stamp(0, it, 0);
      // ---- P[dst, offA + 32*qq ..] (rows of a warp usually share the destination: broadcast loads) stays in flight while we wait
      //      for the gathered source row in the staging tile (column atom qq); rare edge types add their precomputed gaussian block
      float4 av[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        av[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && !(dbg & 4)) av[c] = __ldg(reinterpret_cast<const float4*>(P + (size_t)d0 * TD_NPROJ + m.offA + 32 * qq + 4 * c));
      }
      f2 x[16];
      mbar_wait(bar(B_S_FULL), ph);
      if (rwp == 0) stamp(0, it, 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = lds128(sS + (uint32_t)qq * kAtom + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4));
        x[2 * c] = add2(pk2(v.x, v.y), pk2(av[c].x, av[c].y)); x[2 * c + 1] = add2(pk2(v.z, v.w), pk2(av[c].z, av[c].w));
      }
      if (valid && t0 != 3 && !(dbg & 8)) {
        const float* tr = tslow + (size_t)((blockIdx.x + it * (long long)gridDim.x) * 128 + r) * TD_H + 32 * qq;
        float tv[4][8];
#pragma unroll
        for (int c = 0; c < 4; ++c) ldg256_stream(tr + 8 * c, tv[c]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 4; ++i) x[4 * c + i] = add2(x[4 * c + i], pk2(tv[c][2 * i], tv[c][2 * i + 1]));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_S_EMPTY));
      // ---- + gaussian/type block from the tensor core
      mbar_wait(bar(B_DPRE_FULL0 + (int)ph), (uint32_t)((it >> 1) & 1));
      if (rwp == 0) stamp(0, it, 2);
      tc_fence_after();
      {
        uint32_t v0[16], v1[16];
        const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + 256u + ph * 128u + (uint32_t)(32 * qq);
        tmem_ld16_nowait(ta, v0);
        tmem_ld16_nowait(ta + 16u, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          x[i] = add2(x[i], pk2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1])));
          x[8 + i] = add2(x[8 + i], pk2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1])));
        }
      }
      tc_fence_before();
      // ---- gaussians of the NEXT tile now (the small MMA and its round trip overlap this tile's LayerNorm), metadata two ahead
      if (it + 1 < my_tiles) write_g(s1, t1, dist1);
      s0 = s1; t0 = t1; d0 = d1; dist0 = dist1;
      load_md(it + 2, s1, t1, d1, dist1);
      // the next tile's destination row quarter (and precomputed gaussian row of the rare types) -> L1 while this tile is normalised
      if (s0 >= 0) {
        prefetch_l1(P + (size_t)d0 * TD_NPROJ + m.offA + 32 * qq);
        if (t0 != 3) prefetch_l2(tslow + (size_t)((blockIdx.x + (it + 1) * (long long)gridDim.x) * 128 + r) * TD_H + 32 * qq);
      }
      if (rwp == 0) stamp(0, it, 3);
      // ---- LayerNorm over the 128 features of the row: 4 threads (feature quarters) exchange partial sums through smem
      f2 sa = add2(x[0], x[1]), sb = add2(x[2], x[3]), sc = add2(x[4], x[5]), sd = add2(x[6], x[7]);
      sa = add2(sa, add2(x[8], x[9])); sb = add2(sb, add2(x[10], x[11])); sc = add2(sc, add2(x[12], x[13])); sd = add2(sd, add2(x[14], x[15]));
      float p0, p1;
      upk2(add2(add2(sa, sb), add2(sc, sd)), p0, p1);
      sts32f(xslot + (uint32_t)qq * 512u, p0 + p1);
      named_bar_sync(1 + q, 128);
      const float mean = ((lds32f(xslot) + lds32f(xslot + 512u)) + (lds32f(xslot + 1024u) + lds32f(xslot + 1536u))) * (1.0f / 128.0f);
      named_bar_sync(1 + q, 128);                 // everybody has read the sums before the slots are reused
      const f2 mean2 = pk2(mean, mean);
      f2 qa = pk2(0.f, 0.f), qb = qa, qc = qa, qd = qa;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        x[i] = sub2(x[i], mean2); x[i + 1] = sub2(x[i + 1], mean2); x[i + 2] = sub2(x[i + 2], mean2); x[i + 3] = sub2(x[i + 3], mean2);
        qa = fma2(x[i], x[i], qa); qb = fma2(x[i + 1], x[i + 1], qb); qc = fma2(x[i + 2], x[i + 2], qc); qd = fma2(x[i + 3], x[i + 3], qd);
      }
      upk2(add2(add2(qa, qb), add2(qc, qd)), p0, p1);
      sts32f(xslot + (uint32_t)qq * 512u, p0 + p1);
      named_bar_sync(1 + q, 128);
      const float var = ((lds32f(xslot) + lds32f(xslot + 512u)) + (lds32f(xslot + 1024u) + lds32f(xslot + 1536u))) * (1.0f / 128.0f);
      named_bar_sync(1 + q, 128);                 // slots are rewritten early in the next tile
      const float rstd = rsqrtf(var + 1e-5f);
      // ---- affine + ReLU, bf16 split, store into the activation tile: features 32*qq + 8*c .. -> K-half qq/2, chunk 4*(qq&1)+c
      if (rwp == 0) stamp(0, it, 4);
      mbar_wait(bar(B_A_EMPTY), ph ^ 1u);
      if (rwp == 0) stamp(0, it, 5);
      if (valid) {
        {
          // one copy of this stage for all feature quarters (g / b through indexed constant loads): per-quarter instantiations
          // cost more in instruction-cache misses and spills than they save
          const f2 rstd2 = pk2(rstd, rstd);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int f = 32 * qq + 8 * c + 2 * i;
              const f2 a2 = mul2(rstd2, pk2(lp.g[f], lp.g[f + 1]));
              upk2(fma2(x[4 * c + i], a2, pk2(lp.b[f], lp.b[f + 1])), y[2 * i], y[2 * i + 1]);
              y[2 * i] = fmaxf(y[2 * i], 0.f);
              y[2 * i + 1] = fmaxf(y[2 * i + 1], 0.f);
            }
            const uint32_t addr = sA + (uint32_t)(qq >> 1) * kAtom + (uint32_t)r * 128u + (uint32_t)(((4 * (qq & 1) + c) ^ (r & 7)) << 4);
            split8_store(addr, addr + kPiece, y);
          }
        }
      } else {                                      // absent edge / row beyond the end: zero activation row
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t addr = sA + (uint32_t)(qq >> 1) * kAtom + (uint32_t)r * 128u + (uint32_t)(((4 * (qq & 1) + c) ^ (r & 7)) << 4);
          sts128(addr, 0u, 0u, 0u, 0u);
          sts128(addr + kPiece, 0u, 0u, 0u, 0u);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_A_FULL));
      if (rwp == 0) stamp(0, it, 6);
    }
  } else if (warp >= kGatherWarp0) {
    // ================================================================= gather warps (lane = 4 features), 32 rows each; the last one
    //                                                                   also issues the MMAs (one thread) between its copies
    reg_dec<40>();     // register budget: 256*72 (epilogue, launch value) + 128*40 (gather / MMA) + 512*80 (rows) = 64512 = 896 x 72
    const int gw = warp - kGatherWarp0;
    const bool mma_warp = warp == kMmaWarp;
    const int atom = lane >> 3, ch = lane & 7;
    // source node of row 32*gw + lane of tile t
    auto load_md = [&](long long t) -> int {
      int s_ = -1;
      if (t < my_tiles) {
        const long long idx = (blockIdx.x + t * (long long)gridDim.x) * 128 + 32 * gw + lane;
        if (idx < n_rows) {
          int j;
          const unsigned a = row_dst(idx, j);
          const int dst = row_nodes ? row_nodes[a] : (int)a;
          s_ = src[(size_t)dst * k + j];
        }
      }
      return s_;
    };
    // Dpre[t&1] = G(t) . Tab3^T   (K = 32: two K=16 instructions per product term)
    auto issue_pre = [&](long long t) {
      stamp(2, t, 0);
      mbar_wait(bar(B_G_FULL), (uint32_t)(t & 1));
      stamp(2, t, 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d_addr = tmem_base + 256u + (uint32_t)(t & 1) * 128u;
        uint32_t accum = 0;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const int pa_ = (term == 2) ? 1 : 0, pb_ = (term == 1) ? 1 : 0;      // a1b1, a1b2, a2b1
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            umma_bf16(d_addr, desc_sw64(sG + pa_ * kGPiece + kk * 32), desc_sw64(sT + pb_ * kGPiece + kk * 32), kIdesc, accum);
            accum = 1;
          }
        }
        umma_commit(bar(B_DPRE_FULL0 + (int)(t & 1)));
      }
      __syncwarp();
    };
    // D[t&1] = A(t) . W2^T
    auto issue_main = [&](long long t) {
      const uint32_t ph = (uint32_t)(t & 1), ph2 = (uint32_t)((t >> 1) & 1);
      stamp(2, t, 2);
      mbar_wait_relaxed(bar(B_D_EMPTY0 + (int)ph), ph2 ^ 1u);
      stamp(2, t, 3);
      mbar_wait(bar(B_A_FULL), ph);
      stamp(2, t, 4);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d_addr = tmem_base + ph * 128u;
        uint32_t accum = 0;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const int pa_ = (term == 2) ? 1 : 0, pb_ = (term == 1) ? 1 : 0;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t koff = (kk >> 2) * kAtom + (kk & 3) * 32, woff = (kk >> 2) * kWAtom + (kk & 3) * 32;
            umma_bf16(d_addr, desc_sw128(sA + pa_ * kPiece + koff), desc_sw128(sW + pb_ * kWPiece + woff), kIdescMain, accum);
            accum = 1;
          }
        }
        umma_commit(bar(B_A_EMPTY));
        umma_commit(bar(B_D_FULL0 + (int)ph));
      }
      __syncwarp();
      stamp(2, t, 5);
    };
    int s0 = load_md(0);
    // iteration `it`: copy S(it) (overlaps the row threads' work on tile it-1), then Dpre(it), then the main MMA of tile it-1
    // (its operands become ready when the row threads finish tile it-1, i.e. just before they need S(it))
    for (long long it = 0; it <= my_tiles; ++it) {
      if (it < my_tiles) {
        const int s1 = load_md(it + 1);                      // next tile's metadata lands while this tile's rows are copied
        if (gw == 0) stamp(1, it, 0);
        mbar_wait(bar(B_S_EMPTY), (uint32_t)((it & 1) ^ 1));
        if (gw == 0) stamp(1, it, 1);
        // ---- P[src_row, offB + 4*lane ..] -> S, 32 rows x 512 B per warp, asynchronously (no registers, L2 -> shared)
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const int row = 32 * gw + rr;
          const int sr = __shfl_sync(0xffffffffu, s0, rr);
          const uint32_t dsta = sS + (uint32_t)atom * kAtom + (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
          if (sr >= 0 && !(dbg & 2)) cp_async16(dsta, P + (size_t)sr * TD_NPROJ + m.offB + 4 * lane);
          else sts128f(dsta, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        cp_async_wait_all();
        if (gw == 0) stamp(1, it, 2);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_S_FULL));
        if (gw == 0) stamp(1, it, 3);
        s0 = s1;
        if (mma_warp) issue_pre(it);
      }
      if (mma_warp && it >= 1) issue_main(it - 1);
    }
  } else {
    // ================================================================= epilogue: warp w <-> TMEM lanes 32 (w%4) .., columns 64 (w/4) ..
    const int eq = warp & 3;
    const int HALF = warp >> 2;                    // one code copy for both column halves (b2 through indexed constant loads)
    {
      for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = blockIdx.x + it * gridDim.x;
        const uint32_t ph = (uint32_t)(it & 1), ph2 = (uint32_t)((it >> 1) & 1);
        if (warp == 0) stamp(3, it, 0);
        const uint32_t tbase = tmem_base + ((uint32_t)(eq * 32) << 16) + ph * 128u + (uint32_t)(64 * HALF);
        // fused aggregation (value launch, k == 32): everything that does not depend on the accumulator is fetched before waiting for it
        const bool do_agg = NOUT == 128 && qnode == nullptr && agg.logits != nullptr;
        const bool key_sm = NOUT == 128 && qnode != nullptr && agg.key_softmax;    // key launch, k == 32: softmax in this epilogue
        const long long idx = tile * 128 + eq * 32 + lane;
        const long long dslot = tile * 4 + eq;                          // k == 32: destination index of this warp's 32 rows
        const bool active = do_agg && dslot * 32 < n_rows;              // warp-uniform
        const long long dnode = (active && row_nodes) ? row_nodes[dslot] : dslot;
        float w[8], hin[4], ew = 0.f;
        bool valid_e = false;
        int dst = 0;
        if (do_agg) {
          // attention weights alpha * e_w of this destination's 32 edges, heads 8 HALF .. (written by the key launch's epilogue)
          if (active) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(agg.logits + (size_t)idx * TD_HEADS + 8 * HALF + 4 * i));
              w[4 * i] = t4.x; w[4 * i + 1] = t4.y; w[4 * i + 2] = t4.z; w[4 * i + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) hin[j] = (active && lane < 16) ? agg.h[(size_t)dnode * TD_H + 64 * HALF + 16 * j + lane] : 0.0f;
          // next tile: weights and destination row -> L1
          const long long nslot = dslot + 4 * (long long)gridDim.x;
          if (it + 1 < my_tiles && nslot * 32 < n_rows) {
            prefetch_l1(agg.logits + (size_t)(idx + 128 * (long long)gridDim.x) * TD_HEADS + 8 * HALF);
            if (lane < 2) prefetch_l1(agg.h + (size_t)(row_nodes ? row_nodes[nslot] : nslot) * TD_H + 64 * HALF + 32 * lane);
          }
        } else if (NOUT == 128 && qnode != nullptr) {
          if (idx < n_rows) {
            int j;
            const unsigned a = row_dst(idx, j);
            dst = row_nodes ? row_nodes[a] : (int)a;
            if (key_sm) {
              const size_t e = (size_t)dst * k + j;
              valid_e = src[e] >= 0;
              ew = agg.e_w[e];
            }
          }
          // next tile's query half row (2 lines per destination; the 32 rows of a warp share it when k == 32) -> L1
          const long long nidx = idx + 128 * (long long)gridDim.x;
          if (it + 1 < my_tiles && nidx < n_rows && (lane & 15) == 0) {
            int j;
            const unsigned a = row_dst(nidx, j);
            prefetch_l1(qnode + (size_t)(row_nodes ? row_nodes[a] : (int)a) * TD_H + 64 * HALF + 2 * lane);
          }
        }
        mbar_wait_relaxed(bar(B_D_FULL0 + (int)ph), ph2);
        if (warp == 0) stamp(3, it, 1);
        tc_fence_after();
        if (NOUT == 16) {
          // ---- xv: out[row, 0:16] = D[:, 0:16] + b2   (first column half only)
          if (HALF == 0) {
            uint32_t v[16];
            tmem_ld16(tbase, v);
            if (idx < n_rows && !(dbg & 1)) {
              float* orow = out + (size_t)idx * 16;
              stg256(orow, __uint_as_float(v[0]) + lp.b2[0], __uint_as_float(v[1]) + lp.b2[1], __uint_as_float(v[2]) + lp.b2[2],
                     __uint_as_float(v[3]) + lp.b2[3], __uint_as_float(v[4]) + lp.b2[4], __uint_as_float(v[5]) + lp.b2[5],
                     __uint_as_float(v[6]) + lp.b2[6], __uint_as_float(v[7]) + lp.b2[7]);
              stg256(orow + 8, __uint_as_float(v[8]) + lp.b2[8], __uint_as_float(v[9]) + lp.b2[9], __uint_as_float(v[10]) + lp.b2[10],
                     __uint_as_float(v[11]) + lp.b2[11], __uint_as_float(v[12]) + lp.b2[12], __uint_as_float(v[13]) + lp.b2[13],
                     __uint_as_float(v[14]) + lp.b2[14], __uint_as_float(v[15]) + lp.b2[15]);
            }
          }
        } else if (do_agg) {
          // ---- value MLP with the attention aggregation fused in: this warp's 32 rows are the edges of destination 4*tile + eq
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cb = 64 * HALF;
            const int c0 = 16 * j;
            uint32_t v[16];
            tmem_ld16(tbase + (uint32_t)c0, v);
            float t[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float wh = w[c0 / 8 + i / 4];
              upk2(mul2(add2(pk2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                             pk2(lp.b2[cb + c0 + 2 * i], lp.b2[cb + c0 + 2 * i + 1])), pk2(wh, wh)), t[2 * i], t[2 * i + 1]);
            }
            const float tot = warp_transpose_reduce<16, false>(t, lane);
            if (active && lane < 16 && !(dbg & 1)) agg.h[(size_t)dnode * TD_H + cb + c0 + lane] = hin[j] + tot;
          }
        } else if (qnode == nullptr) {
          // ---- value MLPs: out[row, 64 HALF .. + 64] = D + b2
          float* orow = out + (size_t)idx * 128 + 64 * HALF;
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tbase + (uint32_t)c0, v);
            if (idx < n_rows && !(dbg & 1)) {
#pragma unroll
              for (int c = 0; c < 32; c += 8) {
                const float* bb = lp.b2 + 64 * HALF + c0 + c;
                stg256(orow + c0 + c, __uint_as_float(v[c]) + bb[0], __uint_as_float(v[c + 1]) + bb[1], __uint_as_float(v[c + 2]) + bb[2],
                       __uint_as_float(v[c + 3]) + bb[3], __uint_as_float(v[c + 4]) + bb[4], __uint_as_float(v[c + 5]) + bb[5],
                       __uint_as_float(v[c + 6]) + bb[6], __uint_as_float(v[c + 7]) + bb[7]);
              }
            }
          }
        } else {
          // ---- key MLPs: the keys never leave the SM.  out[row, 8 HALF .. + 8] = attention logits sum_d q[dst, 8h+d] k[row, 8h+d] / sqrt(8)
          //      (reference models/uni_transformer.py:73,135); thread = edge row, q row of the destination read as broadcast loads.
          const float* qrow = qnode + (size_t)dst * TD_H + 64 * HALF;
          float lg[8];
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) {
            float4 qv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = __ldg(reinterpret_cast<const float4*>(qrow + c0 + 4 * i));
            uint32_t v[16];
            tmem_ld16(tbase + (uint32_t)c0, v);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float* bb = lp.b2 + 64 * HALF + c0 + 8 * hh;
              const float4 qa = qv[2 * hh], qb = qv[2 * hh + 1];
              float sacc = (__uint_as_float(v[8 * hh]) + bb[0]) * qa.x;
              sacc = fmaf(__uint_as_float(v[8 * hh + 1]) + bb[1], qa.y, sacc);
              sacc = fmaf(__uint_as_float(v[8 * hh + 2]) + bb[2], qa.z, sacc);
              sacc = fmaf(__uint_as_float(v[8 * hh + 3]) + bb[3], qa.w, sacc);
              sacc = fmaf(__uint_as_float(v[8 * hh + 4]) + bb[4], qb.x, sacc);
              sacc = fmaf(__uint_as_float(v[8 * hh + 5]) + bb[5], qb.y, sacc);
              sacc = fmaf(__uint_as_float(v[8 * hh + 6]) + bb[6], qb.z, sacc);
              sacc = fmaf(__uint_as_float(v[8 * hh + 7]) + bb[7], qb.w, sacc);
              lg[c0 / 8 + hh] = sacc * 0.35355339059327373f;          // 1/sqrt(8)
            }
          }
          if (key_sm) {
            // softmax over the destination's 32 edges (= this warp's rows) for this warp's 8 heads, times the edge gate: the value
            // launch's epilogue only has to weight and sum.  Head hh's max / sum end up in the lanes = hh (mod 8), then are broadcast.
            float tmp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = lg[i] = valid_e ? lg[i] * 1.4426950408889634f : -INFINITY;
            const float mx_mine = warp_transpose_reduce<8, true>(tmp, lane);
#pragma unroll
            for (int hh = 0; hh < 8; ++hh) {
              const float mx = __shfl_sync(0xffffffffu, mx_mine, hh);
              lg[hh] = valid_e ? ex2_approx(lg[hh] - mx) : 0.0f;
              tmp[hh] = lg[hh];
            }
            const float l_mine = warp_transpose_reduce<8, false>(tmp, lane);
            const float inv_mine = l_mine > 0.0f ? 1.0f / l_mine : 0.0f;
#pragma unroll
            for (int hh = 0; hh < 8; ++hh) lg[hh] = lg[hh] * ew * __shfl_sync(0xffffffffu, inv_mine, hh);      // alpha * e_w
          }
          if (idx < n_rows && !(dbg & 1))
            stg256(out + (size_t)idx * TD_HEADS + 8 * HALF, lg[0], lg[1], lg[2], lg[3], lg[4], lg[5], lg[6], lg[7]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_D_EMPTY0 + (int)ph));
        if (warp == 0) stamp(3, it, 2);
      }
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Pre-pass for the rare edge types: tslow[row] = tab[type][20] + sum_j g_j(dist) * tab[type][j] for rows with type != 3 (6-7 % of the
// rows: every edge that touches a ligand atom).  Row-indexed buffer, only those rows are written.  No shared memory is used, so the
// 43 KB table stays L1-resident.  One warp per such row; lanes 0..19 evaluate its gaussians.
//   slow_list != NULL: rows are slots (x2h launches); the list of type != 3 slots was compacted by edge_const_kernel.
//   slow_list == NULL: scan all n_rows rows of a destination subset (h2x launches: ligand destinations, every row is of type 0 / 2).
__global__ void __launch_bounds__(256)
edge_slow_kernel(const int* __restrict__ src, const unsigned char* __restrict__ etype, const float* __restrict__ dist_arr,
                 const int* __restrict__ row_nodes, long long n_rows, int k, const int* __restrict__ slow_list, const int* __restrict__ n_slow,
                 const int* __restrict__ d_n_dst, const float* __restrict__ tab, const float* __restrict__ offsets, float coeff, float* __restrict__ tslow) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * 8;
  const float mu = offsets[lane < TD_NG ? lane : 0];
  if (d_n_dst) n_rows = (long long)(*d_n_dst) * k;
  const long long n_items = slow_list ? (long long)*n_slow : n_rows;
  // 32 items per warp iteration: every lane fetches the metadata of one row (one latency for 32 rows), rows are then processed in turn
  for (long long i0 = warp0 * 32; i0 < n_items; i0 += nwarps * 32) {
    const long long i = i0 + lane;
    long long row = -1;
    int ty = 3;
    float dist = 0.f;
    if (i < n_items) {
      size_t e;
      if (slow_list) {
        e = (size_t)slow_list[i];
        row = (long long)e;
      } else {
        row = i;
        const unsigned a = (unsigned)row / (unsigned)k;
        e = (size_t)row_nodes[a] * k + (row - (long long)a * k);
      }
      ty = etype[e];
      dist = dist_arr[e];
      if (src[e] < 0 || ty == 3) row = -1;
    }
    // The kernel is bound by L1 bandwidth (21 table rows x 512 B per edge row), so edge rows of the same type are processed four at
    // a time in registers: one table-row load feeds four accumulators.
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      unsigned mask = __ballot_sync(0xffffffffu, row >= 0 && ty == t);
      const float* tb = tab + (size_t)t * TD_TAB * TD_H + 4 * lane;
      while (mask) {
        int rr[4];
        long long rrow[4];
        float gj[4];
        float4 v[4];
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(tb + TD_NG * TD_H));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          rr[u] = mask ? __ffs(mask) - 1 : -1;
          if (mask) mask &= mask - 1;
          const int srcl = rr[u] >= 0 ? rr[u] : 0;
          rrow[u] = __shfl_sync(0xffffffffu, row, srcl);
          const float tmu = __shfl_sync(0xffffffffu, dist, srcl) - mu;
          gj[u] = expf(coeff * (tmu * tmu));
          v[u] = c0;
        }
#pragma unroll 4
        for (int jj = 0; jj < TD_NG; ++jj) {
          const float4 cj = __ldg(reinterpret_cast<const float4*>(tb + jj * TD_H));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float g = __shfl_sync(0xffffffffu, gj[u], jj);
            v[u].x = fmaf(g, cj.x, v[u].x); v[u].y = fmaf(g, cj.y, v[u].y); v[u].z = fmaf(g, cj.z, v[u].z); v[u].w = fmaf(g, cj.w, v[u].w);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rr[u] >= 0) *reinterpret_cast<float4*>(tslow + (size_t)rrow[u] * TD_H + 4 * lane) = v[u];
      }
    }
  }
}


// =====================================================================================================================
// EXPERIMENTAL (off unless TDIFF_SLOW_TC=1; written at the end of round 1 without GPU time left to validate it -- see
// DESIGN.md 6.1): the rare-type gaussian block on tensor cores.  Same arithmetic as the Dpre MMA of edge_mlp_v3_kernel
// (G[128x32] . Tab_t^T, bf16 2-piece split, 3 products), but for tiles of 128 rows of ONE rare type t in {0,1,2} taken from
// per-type slot lists, written to the row-indexed buffer edge_slow_kernel fills today.  Sequential phases per tile
// (gaussians -> MMA -> TMEM read + store); 64 KB shared memory and 128 TMEM columns per CTA, so 3 CTAs share an SM.
// =====================================================================================================================
namespace v3 {
constexpr int kPreThreads = 160;                       // warps 0-3: one thread per tile row; warp 4: MMA issuer + TMEM owner
constexpr int oPT = 0, oPG = oPT + 3 * 2 * kGPiece, oPBar = oPG + 2 * kGPiece, kPreSmem = oPBar + 32;
}

// bucket the compact slow-slot list by edge type (order inside a bucket is irrelevant)
__global__ void slow_bucket_kernel(const int* __restrict__ slow_list, const int* __restrict__ n_slow, const unsigned char* __restrict__ etype,
                                   int* __restrict__ type_list, long long cap, int* __restrict__ n_type) {
  const long long n = *n_slow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int e = slow_list[i];
    const int t = etype[e];
    if (t < 3) type_list[(size_t)t * cap + atomicAdd(&n_type[t], 1)] = e;
  }
}
void td_launch_slow_bucket(const int* slow_list, const int* n_slow, const unsigned char* etype, int* type_list, long long cap, int* n_type,
                           int sm_count, cudaStream_t st) {
  cudaMemsetAsync(n_type, 0, 3 * sizeof(int), st);
  slow_bucket_kernel<<<sm_count * 4, 256, 0, st>>>(slow_list, n_slow, etype, type_list, cap, n_type);
}

__global__ void __launch_bounds__(v3::kPreThreads)
edge_pre_tc_kernel(const int* __restrict__ type_list, long long cap, const int* __restrict__ n_type, int type_mask,
                   const float* __restrict__ dist_arr, int k, const int* __restrict__ node_rank,
                   const unsigned char* __restrict__ tab012_img /* 3 types x 3 pieces x 8 KB */, const float* __restrict__ offsets, float coeff,
                   float* __restrict__ tslow) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sT = sbase + oPT, sG = sbase + oPG, sBarA = sbase + oPBar;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem_raw + oPBar + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((sbase & 1023u) != 0) __trap();
  // tables of the types this launch can meet: pieces 0 and 1 of each image
  for (int t = 0; t < 3; ++t) {
    if (!((type_mask >> t) & 1)) continue;
    for (int i = tid; i < 2 * kGPiece / 16; i += kPreThreads) {
      const uint4 v = reinterpret_cast<const uint4*>(tab012_img + (size_t)t * 3 * kGPiece)[i];
      sts128(sT + (uint32_t)t * 2u * kGPiece + 16u * i, v.x, v.y, v.z, v.w);
    }
  }
  if (tid == 0) {
    mbar_init(sBarA, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(s_tmem), 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  long long tiles_of[3], n_of[3], total = 0;
  for (int t = 0; t < 3; ++t) {
    n_of[t] = ((type_mask >> t) & 1) ? (long long)n_type[t] : 0;
    tiles_of[t] = (n_of[t] + 127) / 128;
    total += tiles_of[t];
  }
  const float coeff2 = coeff * 1.4426950408889634f;
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int t = 0;
    long long lt = tile;
    while (lt >= tiles_of[t]) { lt -= tiles_of[t]; ++t; }
    long long row = -1;
    if (warp < 4) {
      // ---- gaussian row of this thread's edge -> G (K slots: 8*c + i = gaussian 5*c + i, slot 29 = 1), bf16 split
      const int r = tid;
      const long long i = lt * 128 + r;
      float dist = 0.f;
      if (i < n_of[t]) {
        const int e = type_list[(size_t)t * cap + i];
        dist = dist_arr[e];
        row = node_rank ? (long long)node_rank[e / k] * k + (e % k) : (long long)e;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float gv[8];
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          const float d = dist - offsets[5 * c + ii];
          gv[ii] = row >= 0 ? ex2_approx(coeff2 * (d * d)) : 0.0f;
        }
        gv[5] = (row >= 0 && c == 3) ? 1.0f : 0.0f;
        gv[6] = gv[7] = 0.0f;
        const uint32_t a0 = sG + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u + (((uint32_t)c ^ (uint32_t)((r >> 1) & 3)) << 4);
        split8_store(a0, a0 + kGPiece, gv);
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
      tc_fence_after();
      if (lane == 0) {
        uint32_t accum = 0;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const int pa_ = (term == 2) ? 1 : 0, pb_ = (term == 1) ? 1 : 0;      // a1b1, a1b2, a2b1
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            umma_bf16(tmem_base, desc_sw64(sG + pa_ * kGPiece + kk * 32), desc_sw64(sT + (uint32_t)t * 2u * kGPiece + pb_ * kGPiece + kk * 32),
                      kIdesc, accum);
            accum = 1;
          }
        }
        umma_commit(sBarA);
      }
      __syncwarp();
    } else {
      mbar_wait(sBarA, phase);
      tc_fence_after();
      float* orow = tslow + (size_t)(row >= 0 ? row : 0) * TD_H;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (row >= 0) {
#pragma unroll
          for (int c = 0; c < 32; c += 8)
            stg256(orow + c0 + c, __uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]),
                   __uint_as_float(v[c + 4]), __uint_as_float(v[c + 5]), __uint_as_float(v[c + 6]), __uint_as_float(v[c + 7]));
        }
      }
      tc_fence_before();
    }
    phase ^= 1u;
    __syncthreads();                      // G and the accumulator are reused by the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

void td_launch_edge_pre_tc(const TdSlowTc& s, int type_mask, const int* node_rank, const float* dist, int k, const unsigned char* tab012_img,
                           const float* offsets, float coeff, float* tslow, int sm_count, cudaStream_t st) {
  static size_t opted[TD_MAX_DEVICES] = {0};
  td_opt_in_smem(edge_pre_tc_kernel, kPreSmem, opted);
  edge_pre_tc_kernel<<<sm_count * 3, kPreThreads, kPreSmem, st>>>(s.type_list, s.cap, s.n_type, type_mask, dist, k, node_rank, tab012_img, offsets,
                                                                 coeff, tslow);
}

void td_launch_edge_mlp_v3(const float* P, const int* src, const unsigned char* etype, const float* dist, const int* row_nodes, long long n_rows,
                           int k, TdMlp m, const unsigned char* w2_image, const unsigned char* tab3_image, const float* offsets, float coeff,
                           const float* h_ln_g, const float* h_ln_b, const float* h_b2, float* tslow, const int* slow_list, const int* n_slow,
                           const float* qnode, float* out, const float* agg_logits, const float* agg_e_w, float* agg_h, int agg_n_nodes,
                           const int* d_n_dst, int key_softmax, const TdSlowTc* stc, int sm_count, cudaStream_t st) {
  if (n_rows == 0) return;
  LnParams lp;
  memcpy(lp.g, h_ln_g, sizeof(lp.g));
  memcpy(lp.b, h_ln_b, sizeof(lp.b));
  memset(lp.b2, 0, sizeof(lp.b2));
  memcpy(lp.b2, h_b2, sizeof(float) * (size_t)m.nout);
  static size_t opted128[TD_MAX_DEVICES] = {0}, opted16[TD_MAX_DEVICES] = {0};
  td_opt_in_smem(edge_mlp_v3_kernel<128>, kSmem, opted128);
  td_opt_in_smem(edge_mlp_v3_kernel<16>, kSmem, opted16);
  const long long n_tiles = (n_rows + 127) / 128;
  const int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
  // experimental tensor-core pre-pass (TDIFF_SLOW_TC=1): full x2h launches (rows = slots) and h2x launches (rows = ligand rank * k + j)
  const bool pre_tc = stc && stc->type_list && m.tab012_img && d_n_dst == nullptr && (row_nodes == nullptr || stc->node_rank != nullptr);
  if (pre_tc) {
    td_launch_edge_pre_tc(*stc, row_nodes ? 0x5 : 0x7, row_nodes ? stc->node_rank : nullptr, dist, k, m.tab012_img, offsets, coeff, tslow, sm_count, st);
  } else {
    const bool listed = (row_nodes == nullptr) && slow_list;            // x2h: iterate the compacted list; h2x: scan the (few) rows
    long long blocks = (listed || d_n_dst) ? sm_count * 8 : (n_rows + 255) / 256;      // a block covers 256 rows per grid-stride iteration
    if (blocks > sm_count * 8) blocks = sm_count * 8;
    edge_slow_kernel<<<(int)blocks, 256, 0, st>>>(src, etype, dist, row_nodes, n_rows, k, listed ? slow_list : nullptr, n_slow, d_n_dst, m.tab, offsets,
                                                  coeff, tslow);
  }
  AggArgs agg = {agg_logits, agg_e_w, agg_h, agg_n_nodes, (key_softmax && k == 32) ? 1 : 0};
  static int dbg = -1;
  static long long* d_ts = nullptr;
  if (dbg < 0) {
    const char* e = getenv("TDIFF_V3_DBG");
    dbg = e ? atoi(e) : 0;
    if (getenv("TDIFF_V3_TS")) cudaMalloc(&d_ts, 16 * 4 * 8 * 8);
  }
  if (d_ts) cudaMemsetAsync(d_ts, 0, 16 * 4 * 8 * 8, st);
  if (m.nout == 16)
    edge_mlp_v3_kernel<16><<<grid, kThreads, kSmem, st>>>(P, src, etype, dist, row_nodes, n_rows, k, m, w2_image, tab3_image, offsets, coeff, tslow,
                                                         nullptr, out, agg, d_n_dst, dbg, lp, d_ts);
  else
    edge_mlp_v3_kernel<128><<<grid, kThreads, kSmem, st>>>(P, src, etype, dist, row_nodes, n_rows, k, m, w2_image, tab3_image, offsets, coeff, tslow,
                                                          qnode, out, agg, d_n_dst, dbg, lp, d_ts);
  if (d_ts && n_rows > 1000000) {          // dump the timeline of the TDIFF_V3_TS-th big launch, once
    static int seen = 0;
    if (++seen == atoi(getenv("TDIFF_V3_TS"))) {
      long long h[16 * 4 * 8];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, d_ts, sizeof(h), cudaMemcpyDeviceToHost);
      const long long t00 = h[(4 * 4 + 0) * 8 + 0];
      const char* names[4] = {"row", "gather", "mma", "epi"};
      for (int t = 4; t < 12; ++t)
        for (int ro = 0; ro < 4; ++ro) {
          printf("tile %2d %-6s", t, names[ro]);
          for (int ev = 0; ev < 7; ++ev) printf(" %8lld", h[(t * 4 + ro) * 8 + ev] ? h[(t * 4 + ro) * 8 + ev] - t00 : -1LL);
          printf("\n");
        }
      fflush(stdout);
    }
  }
}
