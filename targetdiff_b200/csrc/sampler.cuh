// sampler.cuh -- argument block of the fused step epilogue and the marshalling launchers (sampler.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct TdStepArgs {
  int n_lig, n_classes, t_start, pos_only;
  int* step;                       // device: steps already done in this chain (advanced by the launcher's tail kernel)
  const int* lig_node;             // [Nl] node index of each ligand atom
  const int* lig_graph;            // [Nl] graph id
  const float4* xm_final;          // node array after the last layer (predicted x0 in the centred frame)
  const float* logits;             // [Nl,K]
  const float4* offset;            // [B] pocket centroids
  const float *c0, *ct, *logvar;   // posterior_mean_c0_coef, posterior_mean_ct_coef, posterior_logvar [T]
  const float *sra, *srm1;         // sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod [T] (model_mean_type 'noise')
  int mean_noise;                  // 1: the network output is x_t + predicted noise direction (reference :663-666)
  const float *la_v, *l1ma_v, *lca_v, *l1mca_v;   // log_alphas_v, log_one_minus_alphas_v, and their cumprod versions [T]
  float log_k;                     // float32(np.log(num_classes))
  const float* pos_noise;          // tape [S,Nl,3] or NULL (Philox)
  const float* v_uniform;          // tape [S,Nl,K] or NULL (Philox)
  unsigned long long seed;
  float4* lig_pos;                 // in/out [Nl] centred ligand positions
  int* lig_v;                      // in/out [Nl]
  float* pos_traj;                 // [S,Nl,3] or NULL
  long long* v_traj;               // [S,Nl] or NULL
  float* v0_traj;                  // [S,Nl,K] or NULL
  float* vt_traj;                  // [S,Nl,K] or NULL
};

void td_launch_step_epilogue(const TdStepArgs& A, cudaStream_t st);
void td_launch_segment_mean3(const float* pos, const int* seg_ptr, int n_seg, float4* out, cudaStream_t st);
void td_launch_place_protein(const float* pos, const int* prot_node, const int* prot_graph, const float4* offset, int n, float4* xm0,
                             float4* xm1, cudaStream_t st);
void td_launch_park_ligand(float4* lig_pos, int* lig_v, int n, cudaStream_t st);
void td_launch_set_ligand(const float* pos, const long long* v, const int* lig_graph, const float4* offset, int apply_center, int n,
                          int n_classes, float4* lig_pos, int* lig_v, int* err, cudaStream_t st);
void td_launch_get_ligand(const float4* lig_pos, const int* lig_v, const int* lig_graph, const float4* offset, int add_offset, int n,
                          float* pos, long long* v, cudaStream_t st);
void td_launch_scatter_ligand_pos(const float4* lig_pos, const int* lig_node, int n, float4* xm, cudaStream_t st);
void td_launch_gather_xyz(const float4* xm, const int* idx, int n, float* out, cudaStream_t st);
void td_launch_pack_xyzm(const float* x, const unsigned char* mask, int n, float4* xm, cudaStream_t st);
void td_launch_edge_count_scan(const int* src, int n_nodes, int k, long long* node_off, long long* total, cudaStream_t st);
void td_launch_edge_compact(const int* src, const float* e_w, int n_nodes, int k, const long long* node_off, const long long* total,
                            long long* edge_index, float* ew_out, cudaStream_t st);
