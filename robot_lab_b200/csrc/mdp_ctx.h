// mdp_ctx.h - the opaque context behind RlCtx* and the interface between the two step-kernel translation units of
// libmdpstep.so: csrc/mdp_step.cu (C-ABI, the general kernel: any spec, any strides, env-id lists, ragged tiles) and
// csrc/mdp_step_v2.cu (the cluster kernels of the two launch kinds an env step consists of).
#ifndef RL_MDP_CTX_H_
#define RL_MDP_CTX_H_

#include "mdp_terms.cuh"

struct RlV2State;   // tensor-map cache + launch bookkeeping of the cluster kernels (mdp_step_v2.cu)

struct RlCtx {
  int device;
  int slot;
  RlStepSpec spec;
  int NW;                 // warps per tile of the general kernel (kE = 32 envs per tile is fixed)
  rlk::Layout L;
  rlk::Schedule* sched_dev;
  rlk::Schedule sched;    // host copy of the schedule of the current launch config
  unsigned int* ticket;
  uint32_t* cta_mask;
  float* log_partials;
  int cta_mask_cap;
  RlRewardTerm* adhoc_dev;
  float* action_tab_dev;  // per-column action table [5][A]: scale, offset, clip lo / hi, joint id | target kind << 8 (rl_process_action)
  float* cj_dev;          // per-joint constants [5][J] (default pos / vel, soft limits, velocity limit) for the kernels' records
  int sm_count;
  size_t smem_optin;      // largest dynamic shared memory a CTA may ask for on this device
  int use_pdl;
  long long* dbg;
  int baked;              // index into RL_BAKED_LIST when the spec equals a build-time specialised one, else -1
  RlV2State* v2;          // NULL: the cluster kernels do not apply to this context (generic spec, switched off)
};

// ---- mdp_step_v2.cu ------------------------------------------------------------------------------------------------
enum { RL_V2_PRE = 0, RL_V2_POST = 1 };   // DONES|REWARDS|COMPACT  /  RESET(masked)|COMMAND|OBS

// Creates ctx->v2 when the context's spec is a baked one the cluster kernels cover (and RL_MDPSTEP_V2 != "0").
int rl_v2_create(RlCtx* ctx);
void rl_v2_destroy(RlCtx* ctx);
// Launches the cluster kernel of `kind` when the launch qualifies (full clusters, SoA fields that a tensor map can
// describe, contiguous sensor rows, no env-id list); *handled = false means: use the general kernel. `a` is the
// parameter block fill_args() built for the general kernel.
int rl_v2_try_launch(RlCtx* ctx, const rlk::KArgs& a, int kind, cudaStream_t st, bool* handled);
// what rl_ctx_get_cluster_config reports: (0, 0) = no cluster kernel applies to a launch of num_envs
void rl_v2_config_for(const RlCtx* ctx, int64_t num_envs, int* cluster_size, int* tiles_per_cta, int* warps_per_cta, long long* launches);

#endif  // RL_MDP_CTX_H_
