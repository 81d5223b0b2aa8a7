// mdp_step.cu - fused per-step MDP pipeline for sm_100a (B200).
//
// One launch evaluates, for every env of a CTA's tile: termination terms, all reward terms (+ episode sums
// and per-term step rewards), the velocity-command update, both observation groups and the ordered
// compaction of reset ids. Reference behaviour restated (paths relative to /root/reference, V/ =
// source/robot_lab/robot_lab/tasks/manager_based/locomotion/velocity/): V/mdp/rewards.py:22-687,
// V/mdp/observations.py:17-35, V/mdp/commands.py:22-85, V/velocity_env_cfg.py:106-254,379-664 and the
// IsaacLab manager loops / upstream terms listed in SURVEY.md Appendix A.
//
// Execution model (HBM-bound arithmetic intensity, ~0.5 FLOP/B; no tensor cores - nothing here is a contraction):
//   * thread-per-env: a warp owns a tile of 32 consecutive envs, lane e its env e. All per-env state is read
//     straight from global memory; with the SoA [C][N] layout every warp load is one coalesced 128-byte line;
//   * task-sliced grid: the terms are packed into G groups of similar cost and a CTA = (group, block of tiles).
//     Every warp of a CTA runs the same straight-line code (baked spec: dispatch, weights, index lists are
//     immediates) on a different tile - instruction fetch is shared, 16-32 independent warps per SM hide the
//     dependent-instruction latency, and no SIMT lane repeats another lane's scalar work;
//   * the last CTA to finish a tile block finalises it (ordered reward sum, late terms, command commit,
//     episode length, done masks) and the last block compacts the reset ids in ascending order.
// profiles/r1_*.md documents why (three earlier mappings measured with ncu + clock64 stamps).
//
// Built with -fmad=false on purpose: the reference is eager PyTorch, every op rounds on its own, and not
// contracting a*b+c keeps threshold decisions (contact > 1 N, |cmd| > 0.1, ...) bit-identical.

#include "rl_mdp_step.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <utility>

#include "generated/baked_specs.cuh"

#define RL_SPEC_SLOTS 3
#define RL_PI_F 3.14159265358979323846f
#define RL_LOG_STRIDE 64  // >= RL_MAX_REWARD_TERMS + RL_MAX_DONE_TERMS + 2

__constant__ RlStepSpec c_spec[RL_SPEC_SLOTS];

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      snprintf(g_err, sizeof(g_err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
               __FILE__, __LINE__);                                                               \
      return RL_ECUDA;                                                                            \
    }                                                                                             \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// Per-launch field descriptors. Every per-env field is {ptr, env stride, comp stride} in elements; the kernels are
// thread-per-env, so the SoA layout [C][N] (env stride 1) makes every warp load one coalesced 128-byte line.
// ---------------------------------------------------------------------------------------------------
constexpr int kE = 32;  // envs per tile = lanes per warp: lane e of a warp owns env e of the warp's tile

struct FieldD {
  const void* ptr;
  int es;    // env stride   (elements)
  int cs;    // comp stride  (elements)
};

enum InField {
  IF_ROOT_POS = 0, IF_QUAT, IF_LIN_VEL, IF_ANG_VEL, IF_JPOS, IF_JVEL, IF_JACC, IF_JTAU,
  IF_CAIR, IF_LAIR, IF_CCON, IF_LCON, IF_BPOS, IF_BVEL, IF_RAYPOS, IF_HIST, IF_RAYS,
  IF_CMD, IF_HEAD, IF_TLEFT, IF_MXY, IF_MYAW, IF_EPLEN, IF_SUMS, IF_ACT, IF_PACT, IF_STEPR, IF_COUNT
};

// ---------------------------------------------------------------------------------------------------
// Work schedule ("task-sliced grid"). A task = one reward term (wide body-mask terms in two halves), one
// observation term (the height scan in 64-column chunks), the termination terms, or the command update (which
// also owns the command-dependent observation columns and, on env_ids launches, the manager reset). Tasks are
// packed into G groups of similar cost (longest-processing-time greedy); a CTA = (group, block of tiles) and all
// of its warps run the SAME straight-line code on different 32-env tiles, one lane per env:
//   * no SIMT lane ever repeats another lane's scalar work,
//   * instruction fetch is shared by every warp of the CTA and an SM only ever sees its group's code,
//   * 16-32 independent warps per SM hide the dependent-instruction latency a lone warp cannot.
// constexpr: a baked spec gets its schedule (and straight-line per-group code) at compile time.
// ---------------------------------------------------------------------------------------------------
#define RL_MAX_TASKS 112
enum { TK_REWARD = 0, TK_OBS = 1, TK_DONES = 2, TK_COMMAND = 3 };

struct Task {
  uint8_t kind, a, b, owner;   // REWARD: a = term k, b = half (0/1); OBS: a = group, b = term index
  uint16_t lo, hi;             // REWARD: body-index range [lo, hi); OBS: column range within the term
  uint16_t col0, pad;          // OBS: first column of the term inside the group row
};
struct Schedule {
  int n, groups;
  Task t[RL_MAX_TASKS];
  uint8_t split[RL_MAX_REWARD_TERMS];   // term evaluated as two partial sums (termv[2k] + termv[2k+1])
  uint8_t late[RL_MAX_REWARD_TERMS];    // weight / sums / step reward applied at finalisation (split terms, is_terminated)
};

__host__ __device__ constexpr int popc64(uint64_t m) { int n = 0; while (m) { m &= m - 1; ++n; } return n; }

__host__ __device__ constexpr int reward_cost(const RlRewardTerm& t, const RlStepSpec& s, int nbodies) {
  const int J = popc64(t.joint_mask), F = t.n_idx, T = s.hist_len;
  const int base = 90;  // context (3 quaternion rotations, gate) + weight / sums / step-reward epilogue
  switch (t.type) {
    case RL_REW_JOINT_TORQUES_L2: case RL_REW_JOINT_VEL_L2: case RL_REW_JOINT_ACC_L2: case RL_REW_JOINT_DEVIATION_L1:
    case RL_REW_JOINT_POWER: case RL_REW_STAND_STILL: return base + 5 * J;
    case RL_REW_JOINT_POS_LIMITS: case RL_REW_JOINT_VEL_LIMITS: case RL_REW_JOINT_POS_PENALTY: return base + 8 * J;
    case RL_REW_JOINT_MIRROR: case RL_REW_ACTION_MIRROR: return base + 8 * F;
    case RL_REW_ACTION_SYNC: return base + 30 * F;
    case RL_REW_ACTION_RATE_L2: return base + 5 * s.action.n_actions;
    case RL_REW_UNDESIRED_CONTACTS: case RL_REW_CONTACT_FORCES: return base + nbodies * T * 18;
    case RL_REW_TRACK_LIN_VEL_XY_EXP: case RL_REW_TRACK_ANG_VEL_Z_EXP: case RL_REW_TRACK_ANG_VEL_Z_WORLD_EXP: return base + 40;
    case RL_REW_TRACK_LIN_VEL_XY_YAW_FRAME_EXP: return base + 230;
    case RL_REW_FEET_AIR_TIME: case RL_REW_FEET_CONTACT: case RL_REW_FEET_CONTACT_WITHOUT_CMD: return base + 10 * F;
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: return base + 12 * F;
    case RL_REW_FEET_AIR_TIME_VARIANCE: return base + 40 * F;
    case RL_REW_FEET_GAIT: return base + 230;
    case RL_REW_FEET_STUMBLE: return base + 20 * F;
    case RL_REW_FEET_SLIDE: return base + F * (60 + T * 18);
    case RL_REW_FEET_HEIGHT: return base + 60 * F;
    case RL_REW_FEET_HEIGHT_BODY: return base + 130 * F;
    case RL_REW_FEET_DISTANCE_Y_EXP: case RL_REW_FEET_DISTANCE_XY_EXP: return base + 40 + 60 * F;
    case RL_REW_WHEEL_VEL_PENALTY: return base + 12 * F;
    default: return base;
  }
}

__host__ __device__ constexpr Schedule make_schedule(const RlStepSpec& s, int groups) {
  Schedule sc{};
  int cost[RL_MAX_TASKS] = {};
  int n = 0;
  sc.t[n] = Task{TK_DONES, 0, 0, 0, 0, 0, 0, 0}; cost[n++] = 120;
  {
    int c = 420;  // command update + heading control; plus its observation columns
    for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
      for (int ti = 0; ti < s.obs[g].n_terms; ++ti)
        if (s.obs[g].terms[ti].type == RL_OBS_GENERATED_COMMANDS) c += 30;
    sc.t[n] = Task{TK_COMMAND, 0, 0, 0, 0, 0, 0, 0}; cost[n++] = c;
  }
  for (int k = 0; k < s.num_reward_terms; ++k) {
    const RlRewardTerm& t = s.rewards[k];
    if (t.weight == 0.f) continue;
    if (t.type == RL_REW_IS_TERMINATED) { sc.late[k] = 1; continue; }
    const bool body_sum = (t.type == RL_REW_UNDESIRED_CONTACTS || t.type == RL_REW_CONTACT_FORCES);
    const int nb = popc64(t.body_mask);
    if (body_sum && nb > 8) {
      int seen = 0, mid = 0;  // split after the first nb/2 set bits
      for (int b = 0; b < 64; ++b) if ((t.body_mask >> b) & 1ull) { if (++seen == nb / 2) { mid = b + 1; break; } }
      sc.split[k] = 1; sc.late[k] = 1;
      sc.t[n] = Task{TK_REWARD, (uint8_t)k, 0, 0, 0, (uint16_t)mid, 0, 0}; cost[n++] = reward_cost(t, s, nb / 2);
      sc.t[n] = Task{TK_REWARD, (uint8_t)k, 1, 0, (uint16_t)mid, 64, 0, 0}; cost[n++] = reward_cost(t, s, nb - nb / 2);
    } else {
      sc.t[n] = Task{TK_REWARD, (uint8_t)k, 0, 0, 0, 64, 0, 0}; cost[n++] = reward_cost(t, s, nb);
    }
  }
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
    int col0 = 0;
    for (int ti = 0; ti < s.obs[g].n_terms; ++ti) {
      const RlObsTerm& o = s.obs[g].terms[ti];
      const int per_col = 10 + ((o.has_noise && s.obs[g].enable_corruption) ? 24 : 0);
      const bool needs_ctx = (o.type == RL_OBS_BASE_LIN_VEL || o.type == RL_OBS_BASE_ANG_VEL || o.type == RL_OBS_PROJECTED_GRAVITY);
      if (o.type != RL_OBS_GENERATED_COMMANDS) {  // those columns belong to the command task
        for (int lo = 0; lo < o.dim; lo += 64) {
          const int hi = (lo + 64 < o.dim) ? lo + 64 : o.dim;
          sc.t[n] = Task{TK_OBS, (uint8_t)g, (uint8_t)ti, 0, (uint16_t)lo, (uint16_t)hi, (uint16_t)col0, 0};
          cost[n++] = 20 + (needs_ctx ? 60 : 0) + per_col * (hi - lo);
        }
      }
      col0 += o.dim;
    }
  }
  sc.n = n;
  sc.groups = groups;
  // longest-processing-time greedy over the groups
  int load[64] = {};
  bool done[RL_MAX_TASKS] = {};
  for (int it = 0; it < n; ++it) {
    int best = -1;
    for (int i = 0; i < n; ++i) if (!done[i] && (best < 0 || cost[i] > cost[best])) best = i;
    int w = 0;
    for (int j = 1; j < groups; ++j) if (load[j] < load[w]) w = j;
    sc.t[best].owner = (uint8_t)w; load[w] += cost[best]; done[best] = true;
  }
  return sc;
}

struct KArgs {
  int N;
  int slot;
  uint32_t phases;
  int has_ids;
  int groups, tiles_per_cta, n_blocks;   // grid = n_blocks * groups; CTA = (block qb, group g), warp w -> tile qb*tiles_per_cta + w
  FieldD in[IF_COUNT];
  FieldD is_heading, is_standing;        // uint8
  RlStepOut out;
  RlRandom rnd;
  const int32_t* env_ids;
  const int32_t* n_env_ids;
  const Schedule* sched;   // device copy (generic kernel); baked kernels carry theirs as constexpr data
  // scratch owned by the context
  float* termv;            // [2K][Ncap] weighted term values (raw partial sums for late terms)
  float* cmd_new;          // [3][Ncap]  updated command, committed at finalisation (rewards read the old one)
  int Ncap;
  unsigned int* block_ticket;   // [n_blocks] CTAs of a tile block that have finished
  unsigned int* final_ticket;   // [1] tile blocks that have been finalised
  uint32_t* tile_mask;          // [tiles] done bits of each tile, for the ordered compaction
  float* log_partials;          // [n_blocks][RL_LOG_STRIDE]
  int use_pdl;
  long long* dbg;
  // single-term evaluation (rl_term_eval)
  const RlRewardTerm* adhoc;
  const uint8_t* ext_terminated;
  float* term_out;
};


// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) - counter-based, so noise needs no state and no bytes.
// counter = (global env id, step lo, step hi, stream<<16 | block), key = seed.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

enum { RL_STREAM_COMMAND = 1, RL_STREAM_RESET_COMMAND = 2, RL_STREAM_OBS = 16 };

struct RandState {
  unsigned long long seed, step;   // step already includes the device-side common step counter
  long long env_id_offset;
};
__device__ __noinline__ uint4 rl_philox(const RandState r, long long env, uint32_t stream, uint32_t block) {
  const unsigned long long genv = (unsigned long long)(env + r.env_id_offset);
  uint4 ctr = make_uint4((uint32_t)genv, (uint32_t)r.step, (uint32_t)(r.step >> 32) ^ (uint32_t)(genv >> 32),
                         (stream << 16) | block);
  return philox4x32_10(ctr, make_uint2((uint32_t)r.seed, (uint32_t)(r.seed >> 32)));
}


// ---------------------------------------------------------------------------------------------------
// Math (restates isaaclab.utils.math [IL]: quat_apply, quat_apply_inverse, yaw_quat, wrap_to_pi)
// ---------------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// v - w*t + xyz x t, t = 2*(xyz x v)
__device__ __noinline__ V3 quat_apply_inverse(float w, V3 q, V3 v) {
  V3 t = cross3(q, v);
  t.x *= 2.f; t.y *= 2.f; t.z *= 2.f;
  V3 c = cross3(q, t);
  return V3{(v.x - w * t.x) + c.x, (v.y - w * t.y) + c.y, (v.z - w * t.z) + c.z};
}
__device__ __noinline__ V3 quat_apply(float w, V3 q, V3 v) {
  V3 t = cross3(q, v);
  t.x *= 2.f; t.y *= 2.f; t.z *= 2.f;
  V3 c = cross3(q, t);
  return V3{(v.x + w * t.x) + c.x, (v.y + w * t.y) + c.y, (v.z + w * t.z) + c.z};
}
__device__ __noinline__ float remainder_pos(float a, float b) {  // torch.remainder, b > 0
  float m = fmodf(a, b);
  if (m != 0.f && (m < 0.f)) m += b;
  return m;
}
__device__ __forceinline__ float wrap_to_pi(float a) {
  const float two_pi = 2.f * RL_PI_F;
  float w = remainder_pos(a + RL_PI_F, two_pi);
  return (w == 0.f && a > 0.f) ? RL_PI_F : (w - RL_PI_F);
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
// torch.clamp semantics for NaN are irrelevant here; +-inf behave like fminf/fmaxf.
// transcendental functions behind calls: one code copy each instead of ~100 inlined instructions per use
__device__ __noinline__ float rl_expf(float x) { return expf(x); }
__device__ __noinline__ float rl_tanhf(float x) { return tanhf(x); }
__device__ __noinline__ float rl_atan2f(float y, float x) { return atan2f(y, x); }
__device__ __noinline__ float rl_sinf(float x) { return sinf(x); }
__device__ __noinline__ float rl_cosf(float x) { return cosf(x); }


// ---------------------------------------------------------------------------------------------------
// Spec access policies. DynPolicy interprets the context's spec from __constant__ memory (any task);
// StaticPolicy<B> reads a spec baked in at build time (generated/baked_specs.cuh): every use below is a constant
// expression, so term dispatch, parameters, index lists, loop bounds, shared-memory offsets and the warp
// schedule fold away.
// ---------------------------------------------------------------------------------------------------
struct Scalars {
  int num_joints, num_hist_bodies, hist_len, num_time_bodies, num_asset_bodies, num_rays;
  int num_reward_terms, num_done_terms, max_episode_length, n_actions;
  float step_dt, contact_time_abs_tol;
};
__host__ __device__ constexpr Scalars scalars_of(const RlStepSpec& s) {
  return Scalars{s.num_joints, s.num_hist_bodies, s.hist_len, s.num_time_bodies, s.num_asset_bodies, s.num_rays,
                 s.num_reward_terms, s.num_done_terms, s.max_episode_length, s.action.n_actions,
                 s.step_dt, s.contact_time_abs_tol};
}

__host__ __device__ constexpr int obs_col0(const RlObsGroup& G, int ti) {
  int c = 0;
  for (int i = 0; i < ti; ++i) c += G.terms[i].dim;
  return c;
}

template <int... Is, class F>
__device__ __forceinline__ void static_for(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}

struct DynPolicy {
  static constexpr bool kStatic = false;
  __device__ __forceinline__ static Scalars scalars(const KArgs& a) { return scalars_of(c_spec[a.slot]); }
  __device__ __forceinline__ static const RlCommandCfg& command(const KArgs& a) { return c_spec[a.slot].command; }
  template <int G, class F> __device__ __forceinline__ static void for_tasks(const KArgs& a, int g, F&& f) {
    const RlStepSpec& S = c_spec[a.slot];
    const int n = a.sched->n;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const Task tk = a.sched->t[i];
      if (tk.owner != g) continue;
      if (tk.kind == TK_REWARD) f(tk, S.rewards[tk.a], S.obs[0].terms[0], false);
      else f(tk, S.rewards[0], S.obs[tk.a].terms[tk.b], S.obs[tk.a].enable_corruption != 0);
    }
  }
  // f(term, k, split, late): split = evaluated as two partial sums; late = finished in stage 2
  template <int G, class F> __device__ __forceinline__ static void for_rewards(const KArgs& a, F&& f) {
    const RlStepSpec& S = c_spec[a.slot];
#pragma unroll 1
    for (int k = 0; k < S.num_reward_terms; ++k) f(S.rewards[k], k, a.sched->split[k] != 0, a.sched->late[k] != 0);
  }
  template <class F> __device__ __forceinline__ static void for_dones(const KArgs& a, F&& f) {
    const RlStepSpec& S = c_spec[a.slot];
#pragma unroll 1
    for (int d = 0; d < S.num_done_terms; ++d) f(S.dones[d], d);
  }
  // the command-dependent observation terms: f(term, group, term index, first column, corruption on)
  template <class F> __device__ __forceinline__ static void for_cmd_obs(const KArgs& a, F&& f) {
    const RlStepSpec& S = c_spec[a.slot];
#pragma unroll 1
    for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
      int col0 = 0;
#pragma unroll 1
      for (int ti = 0; ti < S.obs[g].n_terms; ++ti) {
        if (S.obs[g].terms[ti].type == RL_OBS_GENERATED_COMMANDS) f(S.obs[g].terms[ti], g, ti, col0, S.obs[g].enable_corruption != 0);
        col0 += S.obs[g].terms[ti].dim;
      }
    }
  }
  template <class F> __device__ __forceinline__ static void for_joint_consts(const KArgs& a, int j, F&& f) {
    const RlStepSpec& S = c_spec[a.slot];
    f(S.default_joint_pos[j], S.default_joint_vel[j], S.soft_pos_limit_lo[j], S.soft_pos_limit_hi[j], S.soft_vel_limit[j]);
  }
  __device__ __forceinline__ static int obs_dim(const KArgs& a, int g) { return c_spec[a.slot].obs[g].dim; }
};

template <class B>
struct StaticPolicy {
  static constexpr bool kStatic = true;
  __device__ __forceinline__ static constexpr Scalars scalars(const KArgs&) {
    constexpr Scalars s = scalars_of(B::spec);
    return s;
  }
  __device__ __forceinline__ static constexpr RlCommandCfg command(const KArgs&) {
    constexpr RlCommandCfg c = B::spec.command;
    return c;
  }
  template <int G> struct Sched { static constexpr Schedule value = make_schedule(B::spec, G); };
  // all tasks of group GI, in schedule order
  template <int G, int GI, class F> __device__ __forceinline__ static void group_tasks(F&& f) {
    static_for(std::make_integer_sequence<int, Sched<G>::value.n>{}, [&](auto ic) {
      constexpr int i = decltype(ic)::value;
      constexpr Task tk = Sched<G>::value.t[i];
      if constexpr (tk.owner == GI) {
        static constexpr RlRewardTerm rt = B::spec.rewards[tk.kind == TK_REWARD ? tk.a : 0];
        static constexpr RlObsTerm ot = B::spec.obs[tk.kind == TK_OBS ? tk.a : 0].terms[tk.kind == TK_OBS ? tk.b : 0];
        constexpr bool corrupt = B::spec.obs[tk.kind == TK_OBS ? tk.a : 0].enable_corruption != 0;
        f(tk, rt, ot, corrupt);
      }
    });
  }
  // binary search on the (runtime) group id: log2(G) branches straight to the group's contiguous code
  template <int G, int LO, int HI, class F> __device__ __forceinline__ static void dispatch_group(int g, F&& f) {
    if constexpr (HI - LO == 1) {
      group_tasks<G, LO>(f);
    } else {
      constexpr int MID = (LO + HI) / 2;
      if (g < MID) dispatch_group<G, LO, MID>(g, f); else dispatch_group<G, MID, HI>(g, f);
    }
  }
  template <int G, class F> __device__ __forceinline__ static void for_tasks(const KArgs&, int g, F&& f) {
    dispatch_group<G, 0, G>(g, f);
  }
  template <int G, class F> __device__ __forceinline__ static void for_rewards(const KArgs&, F&& f) {
    static_for(std::make_integer_sequence<int, B::spec.num_reward_terms>{}, [&](auto kc) {
      constexpr int k = decltype(kc)::value;
      static constexpr RlRewardTerm t = B::spec.rewards[k];  // static: runtime-indexed lists read it in place
      constexpr bool split = Sched<G>::value.split[k] != 0, late = Sched<G>::value.late[k] != 0;
      f(t, k, split, late);
    });
  }
  template <class F> __device__ __forceinline__ static void for_dones(const KArgs&, F&& f) {
    static_for(std::make_integer_sequence<int, B::spec.num_done_terms>{}, [&](auto dc) {
      constexpr int d = decltype(dc)::value;
      static constexpr RlDoneTerm t = B::spec.dones[d];
      f(t, d);
    });
  }
  template <class F> __device__ __forceinline__ static void for_cmd_obs(const KArgs&, F&& f) {
    static_for(std::make_integer_sequence<int, RL_NUM_OBS_GROUPS * RL_MAX_OBS_TERMS>{}, [&](auto ic) {
      constexpr int g = decltype(ic)::value / RL_MAX_OBS_TERMS, ti = decltype(ic)::value % RL_MAX_OBS_TERMS;
      if constexpr (ti < B::spec.obs[g].n_terms && B::spec.obs[g].terms[ti < B::spec.obs[g].n_terms ? ti : 0].type == RL_OBS_GENERATED_COMMANDS) {
        static constexpr RlObsTerm t = B::spec.obs[g].terms[ti];
        constexpr int col0 = obs_col0(B::spec.obs[g], ti);
        f(t, g, ti, col0, B::spec.obs[g].enable_corruption != 0);
      }
    });
  }
  __device__ __forceinline__ static constexpr int obs_dim(const KArgs&, int g) { return g == 0 ? B::spec.obs[0].dim : B::spec.obs[1].dim; }
};


// ---------------------------------------------------------------------------------------------------
// Per-env context shared by all terms (thread-per-env: lane e = env e of the tile)
// ---------------------------------------------------------------------------------------------------
struct EnvCtx {
  float qw;
  V3 q;          // quaternion xyz
  V3 g;          // projected_gravity_b
  V3 vb, wb;     // root lin / ang velocity in the base frame
  V3 vw, ww;     // world frame
  V3 pos;        // root_pos_w
  float gate;    // clamp(-g.z, 0, 0.7) / 0.7      (V/mdp/rewards.py:34 and 28 other uses)
  float c0, c1, c2;
  float cmd_norm;   // |cmd|_2 over 3 components
  float vxy_norm;   // |v_b.xy|
  bool terminated;
};

// Thread-per-env operand access: field f, component c of this lane's env, straight from global memory. For the
// SoA [C][N] layout consecutive lanes read consecutive addresses (one 128-byte line per warp instruction); any
// other strides work too, slower.
__device__ __forceinline__ float ld_field(const FieldD& fd, long long env, int c) {
  return static_cast<const float*>(fd.ptr)[env * fd.es + (long long)c * fd.cs];
}
__device__ __forceinline__ float ld_field_ro(const FieldD& fd, long long env, int c) {
  return __ldg(static_cast<const float*>(fd.ptr) + env * fd.es + (long long)c * fd.cs);
}
// Every task-phase read goes through ld.global.nc: within one launch no address is re-read after it was written
// (episode sums / metrics are read once by their only writer; the command and the episode length are rewritten
// at finalisation only, after every reader of the tile block is done), so the compiler may hoist and batch ALL
// loads of a group above the stores of its earlier tasks - one memory round trip per CTA instead of one per task.
#define LDF(f, c) ld_field_ro(a.in[(f)], env, (c))
#define JC(arr, j) CS.arr[(j)]   // per-joint constants: lane-uniform index -> constant-bank broadcast

__device__ __forceinline__ bool first_contact(const KArgs& a, const Scalars& S, long long env, int b) {
  const float t = LDF(IF_CCON, b);
  return (t > 0.f) && (t < (S.step_dt + S.contact_time_abs_tol));
}
__device__ __forceinline__ V3 body_vec(const KArgs& a, int f, long long env, int b) {
  return V3{LDF(f, 3 * b + 0), LDF(f, 3 * b + 1), LDF(f, 3 * b + 2)};
}
// max over the history of |F_b| (net_forces_w_history[:, :, b].norm(-1).max(1))
__device__ __forceinline__ float hist_max_norm(const KArgs& a, long long env, int T, int B, int b) {
  float m = 0.f;
  _Pragma("unroll")
  for (int t = 0; t < T; ++t) {
    const int i = (t * B + b) * 3;
    const float fx = LDF(IF_HIST, i), fy = LDF(IF_HIST, i + 1), fz = LDF(IF_HIST, i + 2);
    const float n = sqrtf((fx * fx + fy * fy) + fz * fz);
    m = (t == 0) ? n : fmaxf(m, n);
  }
  return m;
}

// One reward term for env e: raw value (no weight, no dt). [lo, hi) restricts body-mask terms to a body-index
// range (the two halves of a split term add up).
__device__ __forceinline__ float reward_term(const RlRewardTerm& t, const Scalars& S, const RlStepSpec& CS, const KArgs& a,
                                             const long long env, const EnvCtx& c, const int lo, const int hi) {
  const int J = S.num_joints;
  switch (t.type) {
    case RL_REW_IS_TERMINATED: return c.terminated ? 1.f : 0.f;
    case RL_REW_LIN_VEL_Z_L2: return (c.vb.z * c.vb.z) * c.gate;
    case RL_REW_ANG_VEL_XY_L2: return (c.wb.x * c.wb.x + c.wb.y * c.wb.y) * c.gate;
    case RL_REW_FLAT_ORIENTATION_L2: return (c.g.x * c.g.x + c.g.y * c.g.y) * c.gate;
    case RL_REW_BASE_HEIGHT_L2: {
      const float d = c.pos.z - t.p[0];
      return (d * d) * c.gate;
    }
    case RL_REW_UPWARD: {
      const float d = 1.f - c.g.z;
      return d * d;
    }
    case RL_REW_JOINT_TORQUES_L2:
    case RL_REW_JOINT_VEL_L2:
    case RL_REW_JOINT_ACC_L2: {
      const int off = t.type == RL_REW_JOINT_TORQUES_L2 ? IF_JTAU : (t.type == RL_REW_JOINT_VEL_L2 ? IF_JVEL : IF_JACC);
      float s = 0.f;
      _Pragma("unroll")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) { const float v = LDF(off, j); s += v * v; }
      return s;
    }
    case RL_REW_JOINT_DEVIATION_L1: {
      float s = 0.f;
      _Pragma("unroll")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(LDF(IF_JPOS, j) - JC(default_joint_pos, j));
      return s;
    }
    case RL_REW_JOINT_POS_LIMITS: {
      float s = 0.f;
      _Pragma("unroll")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) {
          const float q = LDF(IF_JPOS, j);
          float o = -fminf(q - JC(soft_pos_limit_lo, j), 0.f);
          o += fmaxf(q - JC(soft_pos_limit_hi, j), 0.f);
          s += o;
        }
      return s;
    }
    case RL_REW_JOINT_VEL_LIMITS: {
      float s = 0.f;
      _Pragma("unroll")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += clampf(fabsf(LDF(IF_JVEL, j)) - JC(soft_vel_limit, j) * t.p[0], 0.f, 1.f);
      return s;
    }
    case RL_REW_JOINT_POWER: {
      float s = 0.f;
      _Pragma("unroll")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(LDF(IF_JVEL, j) * LDF(IF_JTAU, j));
      return s;
    }
    case RL_REW_STAND_STILL: {
      float s = 0.f;
      _Pragma("unroll")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(LDF(IF_JPOS, j) - JC(default_joint_pos, j));
      s *= (c.cmd_norm < t.p[0]) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_JOINT_POS_PENALTY: {
      float s = 0.f;
      _Pragma("unroll")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) { const float d = LDF(IF_JPOS, j) - JC(default_joint_pos, j); s += d * d; }
      const float running = sqrtf(s);
      const bool moving = (c.cmd_norm > t.p[2]) || (c.vxy_norm > t.p[1]);
      return (moving ? running : t.p[0] * running) * c.gate;
    }
    case RL_REW_JOINT_MIRROR: {
      float s = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const float d = LDF(IF_JPOS, t.idx_a[i]) - LDF(IF_JPOS, t.idx_b[i]);
        s += d * d;
      }
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_MIRROR: {
      float s = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const float d = fabsf(LDF(IF_ACT, t.idx_a[i])) - fabsf(LDF(IF_ACT, t.idx_b[i]));
        s += d * d;
      }
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_SYNC: {
      float r = 0.f;
      for (int g = 0; g < t.n_idx; ++g) {
        const int start = t.idx_b[g], n = t.idx_c[g];
        if (n < 2) continue;
        float m = 0.f;
        for (int i = 0; i < n; ++i) m += fabsf(LDF(IF_ACT, t.idx_a[start + i]));
        m = m / (float)n;
        float v = 0.f;
        for (int i = 0; i < n; ++i) { const float d = fabsf(LDF(IF_ACT, t.idx_a[start + i])) - m; v += d * d; }
        r += v / (float)n;
      }
      return (r * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_RATE_L2: {
      float s = 0.f;
      _Pragma("unroll")
      for (int ai = 0; ai < S.n_actions; ++ai) { const float d = LDF(IF_ACT, ai) - LDF(IF_PACT, ai); s += d * d; }
      return s;
    }
    case RL_REW_UNDESIRED_CONTACTS: {
      float s = 0.f;
      _Pragma("unroll")
      for (int b = lo; b < S.num_hist_bodies && b < hi; ++b)
        if (((t.body_mask >> b) & 1ull) && (hist_max_norm(a, env, S.hist_len, S.num_hist_bodies, b) > t.p[0])) s += 1.f;
      return s * c.gate;   // gate distributes over the two halves of a split term
    }
    case RL_REW_CONTACT_FORCES: {
      float s = 0.f;
      _Pragma("unroll")
      for (int b = lo; b < S.num_hist_bodies && b < hi; ++b)
        if ((t.body_mask >> b) & 1ull) s += fmaxf(hist_max_norm(a, env, S.hist_len, S.num_hist_bodies, b) - t.p[0], 0.f);
      return s;
    }
    case RL_REW_TRACK_LIN_VEL_XY_EXP: {
      const float dx = c.c0 - c.vb.x, dy = c.c1 - c.vb.y;
      return rl_expf(-(dx * dx + dy * dy) / t.p[0]) * c.gate;
    }
    case RL_REW_TRACK_ANG_VEL_Z_EXP: {
      const float d = c.c2 - c.wb.z;
      return rl_expf(-(d * d) / t.p[0]) * c.gate;
    }
    case RL_REW_TRACK_LIN_VEL_XY_YAW_FRAME_EXP: {
      // yaw_quat [IL] then quat_apply_inverse on the world velocity (V/mdp/rewards.py:60)
      const float yaw = rl_atan2f(2.f * (c.qw * c.q.z + c.q.x * c.q.y), 1.f - 2.f * (c.q.y * c.q.y + c.q.z * c.q.z));
      float yw = rl_cosf(yaw / 2.f), yz = rl_sinf(yaw / 2.f);
      const float nrm = fmaxf(sqrtf(yw * yw + yz * yz), 1e-9f);
      yw = yw / nrm; yz = yz / nrm;
      const V3 v = quat_apply_inverse(yw, V3{0.f, 0.f, yz}, c.vw);
      const float dx = c.c0 - v.x, dy = c.c1 - v.y;
      return rl_expf(-(dx * dx + dy * dy) / t.p[0]) * c.gate;
    }
    case RL_REW_TRACK_ANG_VEL_Z_WORLD_EXP: {
      const float d = c.c2 - c.ww.z;
      return rl_expf(-(d * d) / t.p[0]) * c.gate;
    }
    case RL_REW_FEET_AIR_TIME: {
      float s = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = t.idx_a[i];
        s += (LDF(IF_LAIR, b) - t.p[0]) * (first_contact(a, S, env, b) ? 1.f : 0.f);
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: {
      int n_contact = 0;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) n_contact += (LDF(IF_CCON, t.idx_a[i]) > 0.f) ? 1 : 0;
      const bool single = (n_contact == 1);
      float r = INFINITY;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = t.idx_a[i];
        const float ct = LDF(IF_CCON, b);
        const float mode = (ct > 0.f) ? ct : LDF(IF_CAIR, b);
        r = fminf(r, single ? mode : 0.f);
      }
      r = fminf(r, t.p[0]);
      r *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_AIR_TIME_VARIANCE: {
      // torch.var (unbiased) is a Welford reduction on CPU; keep the same update order.
      float r = 0.f;
      for (int which = 0; which < 2; ++which) {
        const int off = which == 0 ? IF_LAIR : IF_LCON;
        float mean = 0.f, m2 = 0.f;
        _Pragma("unroll")
        for (int i = 0; i < t.n_idx; ++i) {
          const float x = fminf(LDF(off, t.idx_a[i]), 0.5f);
          const float d = x - mean;
          mean += d / (float)(i + 1);
          m2 += d * (x - mean);
        }
        r += m2 / (float)(t.n_idx - 1);
      }
      return r * c.gate;
    }
    case RL_REW_FEET_GAIT: {
      const int f00 = t.idx_a[0], f01 = t.idx_a[1], f10 = t.idx_a[2], f11 = t.idx_a[3];
      const float me2 = t.p[1], sd = t.p[0];
      auto sync = [&](int fa, int fb) {
        const float da = LDF(IF_CAIR, fa) - LDF(IF_CAIR, fb);
        const float dc = LDF(IF_CCON, fa) - LDF(IF_CCON, fb);
        return rl_expf(-(fminf(da * da, me2) + fminf(dc * dc, me2)) / sd);
      };
      auto async = [&](int fa, int fb) {
        const float d0 = LDF(IF_CAIR, fa) - LDF(IF_CCON, fb);
        const float d1 = LDF(IF_CCON, fa) - LDF(IF_CAIR, fb);
        return rl_expf(-(fminf(d0 * d0, me2) + fminf(d1 * d1, me2)) / sd);
      };
      const float sync_r = sync(f00, f01) * sync(f10, f11);
      const float async_r = ((async(f00, f10) * async(f01, f11)) * async(f00, f11)) * async(f10, f01);
      const bool moving = (c.cmd_norm > t.p[3]) || (c.vxy_norm > t.p[2]);
      return (moving ? sync_r * async_r : 0.f) * c.gate;
    }
    case RL_REW_FEET_CONTACT: {
      int n = 0;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) n += first_contact(a, S, env, t.idx_a[i]) ? 1 : 0;
      float r = ((float)n != t.p[0]) ? 1.f : 0.f;
      r *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_CONTACT_WITHOUT_CMD: {
      int n = 0;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) n += first_contact(a, S, env, t.idx_a[i]) ? 1 : 0;
      float r = (float)n;
      r *= (c.cmd_norm < 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_STUMBLE: {
      bool any = false;   // t = 0 is the newest history sample = net_forces_w
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = t.idx_c[i];
        const float fx = LDF(IF_HIST, 3 * b + 0), fy = LDF(IF_HIST, 3 * b + 1), fz = LDF(IF_HIST, 3 * b + 2);
        any = any || (sqrtf(fx * fx + fy * fy) > 4.f * fabsf(fz));
      }
      return (any ? 1.f : 0.f) * c.gate;
    }
    case RL_REW_FEET_SLIDE: {
      float s = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 vw = body_vec(a, IF_BVEL, env, t.idx_b[i]);
        const V3 vb = quat_apply_inverse(c.qw, c.q, V3{vw.x - c.vw.x, vw.y - c.vw.y, vw.z - c.vw.z});
        const float lat = sqrtf(vb.x * vb.x + vb.y * vb.y);
        s += lat * ((hist_max_norm(a, env, S.hist_len, S.num_hist_bodies, t.idx_c[i]) > 1.0f) ? 1.f : 0.f);
      }
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT: {
      float s = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 p = body_vec(a, IF_BPOS, env, t.idx_b[i]);
        const V3 v = body_vec(a, IF_BVEL, env, t.idx_b[i]);
        const float d = p.z - t.p[0];
        s += (d * d) * rl_tanhf(t.p[1] * sqrtf(v.x * v.x + v.y * v.y));
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT_BODY: {
      float s = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec(a, IF_BPOS, env, t.idx_b[i]);
        const V3 vw = body_vec(a, IF_BVEL, env, t.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const V3 vb = quat_apply_inverse(c.qw, c.q, V3{vw.x - c.vw.x, vw.y - c.vw.y, vw.z - c.vw.z});
        const float d = pb.z - t.p[0];
        s += (d * d) * rl_tanhf(t.p[1] * sqrtf(vb.x * vb.x + vb.y * vb.y));
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_DISTANCE_Y_EXP: {
      float s = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec(a, IF_BPOS, env, t.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const float want = (t.p[0] / 2.f) * ((i % 2 == 0) ? 1.f : -1.f);
        const float d = want - pb.y;
        s += d * d;
      }
      return rl_expf(-s / t.p[1]) * c.gate;
    }
    case RL_REW_FEET_DISTANCE_XY_EXP: {
      float s = 0.f;
      for (int i = 0; i < 4; ++i) {
        const V3 pw = body_vec(a, IF_BPOS, env, t.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const float wx = (i < 2) ? (t.p[1] / 2.f) : (-t.p[1] / 2.f);
        const float wy = (i % 2 == 0) ? (t.p[0] / 2.f) : (-t.p[0] / 2.f);
        const float dx = wx - pb.x, dy = wy - pb.y;
        s += dx * dx + dy * dy;
      }
      return rl_expf(-s / t.p[2]) * c.gate;
    }
    case RL_REW_WHEEL_VEL_PENALTY: {
      float run = 0.f, stand = 0.f;
      _Pragma("unroll")
      for (int i = 0; i < t.n_idx; ++i) {
        const float jv = fabsf(LDF(IF_JVEL, t.idx_b[i]));
        const float ta = LDF(IF_CAIR, t.idx_a[i]);
        const bool first_air = (ta > 0.f) && (ta < (S.step_dt + S.contact_time_abs_tol));
        run += (first_air ? 1.f : 0.f) * jv;
        stand += jv;
      }
      const bool moving = (c.cmd_norm > t.p[1]) || (c.vxy_norm > t.p[0]);
      return moving ? run : stand;
    }
    default: return 0.f;
  }
}

__device__ __forceinline__ EnvCtx make_ctx(const KArgs& a, const long long env) {
  EnvCtx c;
  c.qw = LDF(IF_QUAT, 0);
  c.q = V3{LDF(IF_QUAT, 1), LDF(IF_QUAT, 2), LDF(IF_QUAT, 3)};
  c.pos = V3{LDF(IF_ROOT_POS, 0), LDF(IF_ROOT_POS, 1), LDF(IF_ROOT_POS, 2)};
  c.vw = V3{LDF(IF_LIN_VEL, 0), LDF(IF_LIN_VEL, 1), LDF(IF_LIN_VEL, 2)};
  c.ww = V3{LDF(IF_ANG_VEL, 0), LDF(IF_ANG_VEL, 1), LDF(IF_ANG_VEL, 2)};
  c.g = quat_apply_inverse(c.qw, c.q, V3{0.f, 0.f, -1.f});
  c.vb = quat_apply_inverse(c.qw, c.q, c.vw);
  c.wb = quat_apply_inverse(c.qw, c.q, c.ww);
  c.gate = clampf(-c.g.z, 0.f, 0.7f) / 0.7f;
  c.c0 = LDF(IF_CMD, 0); c.c1 = LDF(IF_CMD, 1); c.c2 = LDF(IF_CMD, 2);
  c.cmd_norm = sqrtf((c.c0 * c.c0 + c.c1 * c.c1) + c.c2 * c.c2);
  c.vxy_norm = sqrtf(c.vb.x * c.vb.x + c.vb.y * c.vb.y);
  c.terminated = false;
  return c;
}



// ---------------------------------------------------------------------------------------------------
// Stores of per-env results (thread-per-env; SoA targets coalesce)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_field(const FieldD& fd, long long env, int c, float v) {
  static_cast<float*>(const_cast<void*>(fd.ptr))[env * fd.es + (long long)c * fd.cs] = v;
}
#define STF(f, c, v) st_field(a.in[(f)], env, (c), (v))
__device__ __forceinline__ int ld_u8(const FieldD& fd, long long env) {
  return (int)static_cast<const uint8_t*>(fd.ptr)[env * fd.es];
}
__device__ __forceinline__ void st_u8(const FieldD& fd, long long env, int v) {
  static_cast<uint8_t*>(const_cast<void*>(fd.ptr))[env * fd.es] = (uint8_t)v;
}
__device__ __forceinline__ int ld_eplen(const KArgs& a, long long env) {
  return static_cast<const int32_t*>(a.in[IF_EPLEN].ptr)[env * a.in[IF_EPLEN].es];
}

// Termination terms of one env: TerminationManager.compute [IL]. bits | terminated << 8 | truncated << 9
template <class P>
__device__ __forceinline__ uint32_t eval_dones(const KArgs& a, const Scalars& S, const long long env, const EnvCtx& c) {
  uint32_t bits = 0, term = 0, trunc = 0;
  const int eplen = ld_eplen(a, env) + 1;   // episode_length_buf += 1 precedes the termination terms
  P::for_dones(a, [&](const RlDoneTerm& t, int d) {
    int fired = 0;
    if (t.type == RL_DONE_TIME_OUT) {
      fired = eplen >= S.max_episode_length;
    } else if (t.type == RL_DONE_TERRAIN_OUT_OF_BOUNDS) {
      fired = (t.p[2] != 0.f) && ((fabsf(c.pos.x) > t.p[0]) || (fabsf(c.pos.y) > t.p[1]));
    } else if (t.type == RL_DONE_ILLEGAL_CONTACT) {
      _Pragma("unroll")
      for (int b = 0; b < S.num_hist_bodies; ++b)
        if (((t.body_mask >> b) & 1ull) && (hist_max_norm(a, env, S.hist_len, S.num_hist_bodies, b) > t.p[0])) fired = 1;
    }
    if (fired) { bits |= 1u << d; if (t.time_out) trunc = 1; else term = 1; }
  });
  return bits | (term << 8) | (trunc << 9);
}

// Columns [lo, hi) of one observation term for the warp's 32 envs: ObservationManager.compute_group [IL]
// (clone -> +noise -> clip -> scale). Lane e computes its env's values into a [32][33] shared-memory tile
// (conflict-free both ways), then the warp writes row segments of the [N, D] observation rows: consecutive
// lanes hit consecutive addresses instead of 32 partial sectors per store.
__device__ __forceinline__ void obs_task(const KArgs& a, const Scalars& S, const RlStepSpec& CS, const RlObsTerm& t,
                                         const bool corrupt, const RandState rs, const int g, const int ti,
                                         const int col0, const int lo, const int hi, const long long env,
                                         const bool valid, const EnvCtx& c, const float* cmd3, const int eplen_eff,
                                         const bool zero_action, float (*tile)[33]) {
  const int lane = threadIdx.x & 31;
  float* const obs = a.out.obs[g];
  const long long pitch = a.out.obs_pitch[g];
  const int D = g == 0 ? CS.obs[0].dim : CS.obs[1].dim;
  const float* urow = a.rnd.obs_uniforms[g] ? a.rnd.obs_uniforms[g] + env * D + col0 : nullptr;
  const bool noisy = t.has_noise && corrupt;
  for (int c0 = lo; c0 < hi; c0 += 32) {
    const int nc = min(32, hi - c0);
    _Pragma("unroll")
    for (int q = 0; q < 8; ++q) {
      if (q * 4 >= nc) break;
      float u4[4] = {0.f, 0.f, 0.f, 0.f};
      if (noisy && urow == nullptr) {
        const uint4 r = rl_philox(rs, env, RL_STREAM_OBS + g * RL_MAX_OBS_TERMS + ti, (uint32_t)((c0 >> 2) + q));
        u4[0] = u01(r.x); u4[1] = u01(r.y); u4[2] = u01(r.z); u4[3] = u01(r.w);
      }
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        const int jj = q * 4 + r4;
        const int col = c0 + jj;
        if (jj >= nc) break;
        float v;
        switch (t.type) {
          case RL_OBS_BASE_LIN_VEL: v = col == 0 ? c.vb.x : (col == 1 ? c.vb.y : c.vb.z); break;
          case RL_OBS_BASE_ANG_VEL: v = col == 0 ? c.wb.x : (col == 1 ? c.wb.y : c.wb.z); break;
          case RL_OBS_PROJECTED_GRAVITY: v = col == 0 ? c.g.x : (col == 1 ? c.g.y : c.g.z); break;
          case RL_OBS_GENERATED_COMMANDS: v = cmd3[col]; break;
          case RL_OBS_JOINT_POS_REL: v = LDF(IF_JPOS, t.ids[col]) - JC(default_joint_pos, t.ids[col]); break;
          case RL_OBS_JOINT_POS_REL_WITHOUT_WHEEL:
            v = LDF(IF_JPOS, t.ids[col]) - JC(default_joint_pos, t.ids[col]);
            if ((t.zero_mask >> col) & 1ull) v = 0.f;
            break;
          case RL_OBS_JOINT_VEL_REL: v = LDF(IF_JVEL, t.ids[col]) - JC(default_joint_vel, t.ids[col]); break;
          case RL_OBS_LAST_ACTION: v = zero_action ? 0.f : LDF(IF_ACT, col); break;
          case RL_OBS_HEIGHT_SCAN: v = (LDF(IF_RAYPOS, 0) - LDF(IF_RAYS, col)) - t.p[0]; break;
          case RL_OBS_PHASE: {
            const float ph = ((float)eplen_eff * S.step_dt) / t.p[0];
            v = col == 0 ? rl_sinf((2.f * RL_PI_F) * ph) : rl_cosf((2.f * RL_PI_F) * ph);
            break;
          }
          default: v = 0.f;
        }
        if (noisy) {
          const float u = urow ? urow[col] : u4[r4];
          v = (v + u * (t.noise_hi - t.noise_lo)) + t.noise_lo;
        }
        if (t.has_clip) v = clampf(v, t.clip_lo, t.clip_hi);
        if (t.has_scale) v = v * t.scale;
        tile[lane][jj] = v;
      }
    }
    __syncwarp();
    const int env_lo = (int)env, ok = valid ? 1 : 0;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const long long env_r = (long long)__shfl_sync(0xffffffffu, env_lo, r);
      const int ok_r = __shfl_sync(0xffffffffu, ok, r);
      if (ok_r && lane < nc) obs[env_r * pitch + col0 + c0 + lane] = tile[r][lane];
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void cmd_uniforms(const KArgs& a, const RandState rs, long long env, uint32_t stream, float* u) {
  if (a.rnd.cmd_uniforms != nullptr) {
#pragma unroll
    for (int i = 0; i < RL_NUM_CMD_UNIFORMS; ++i) u[i] = a.rnd.cmd_uniforms[(long long)i * a.N + env];
  } else {
    const uint4 r0 = rl_philox(rs, env, stream, 0), r1 = rl_philox(rs, env, stream, 1);
    u[0] = u01(r0.x); u[1] = u01(r0.y); u[2] = u01(r0.z); u[3] = u01(r0.w);
    u[4] = u01(r1.x); u[5] = u01(r1.y); u[6] = u01(r1.z);
  }
}

// ---------------------------------------------------------------------------------------------------
// The fused step kernel. grid = n_blocks x G, CTA = (tile block qb, task group g), WPC warps; warp w owns tile
// qb*WPC + w, lane e its env e. MODE 0 = step, MODE 1 = single-term evaluation (one group, one task).
// ---------------------------------------------------------------------------------------------------
template <class P, int G, int WPC, int MODE>
__global__ void __launch_bounds__(WPC * 32, 1) mdp_step_kernel(const KArgs a) {
  extern __shared__ __align__(16) float s_dyn[];                 // [WPC][32][33] observation transpose tiles
  float (*xpose)[33] = reinterpret_cast<float (*)[33]>(s_dyn) + (threadIdx.x >> 5) * 32;
  __shared__ float s_red[WPC][RL_LOG_STRIDE];
  __shared__ int s_last, s_final;
  const Scalars S = P::scalars(a);
  const RlStepSpec& CS = c_spec[a.slot];   // lane-uniform table lookups (per-joint constants) in both policies
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t ph = a.phases;
#define RL_STAMP(i) do { if (a.dbg != nullptr && tid == 0) a.dbg[(size_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)
  RL_STAMP(0);
  if (a.use_pdl) {
    // launch-latency overlap only: every read below may depend on the predecessor, so wait first
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int n_total = a.has_ids ? *a.n_env_ids : a.N;
  const int g = MODE == 1 ? 0 : (int)(blockIdx.x % G);
  const int qb = MODE == 1 ? (int)blockIdx.x : (int)(blockIdx.x / G);
  const int tile = qb * WPC + warp;
  const int idx = tile * kE + lane;
  const bool valid = idx < n_total;
  const int n_tiles = (n_total + kE - 1) / kE;
  const int n_blocks_live = (n_tiles + WPC - 1) / WPC;   // tile blocks that hold at least one env
  const int K = S.num_reward_terms;
  const bool do_reset = (MODE == 0) && (ph & RL_PHASE_RESET) && a.has_ids;
  if (do_reset && n_total == 0 && blockIdx.x == 0 && tid < RL_LOG_STRIDE) {
    // nothing to reset: the logged scalars are defined as 0
    if (tid < K) { if (a.out.reset_log.episode_sum_mean) a.out.reset_log.episode_sum_mean[tid] = 0.f; }
    else if (tid < K + RL_MAX_DONE_TERMS) { if (a.out.reset_log.done_term_count) a.out.reset_log.done_term_count[tid - K] = 0.f; }
    else if (tid < K + RL_MAX_DONE_TERMS + 2) { if (a.out.reset_log.metric_mean) a.out.reset_log.metric_mean[tid - K - RL_MAX_DONE_TERMS] = 0.f; }
  }
  if (qb >= n_blocks_live) return;   // whole CTA beyond the data (env_ids launches are sized for the worst case)
  // invalid lanes shadow a valid env (loads stay in bounds) and never store
  const int idx_c = valid ? idx : 0;
  const long long env = n_total > 0 ? (a.has_ids ? (long long)a.env_ids[idx_c] : (long long)idx_c) : 0;
  RandState rs;
  rs.seed = a.rnd.seed;
  rs.step = a.rnd.step + (a.rnd.step_counter ? *a.rnd.step_counter : 0ull);
  rs.env_id_offset = a.rnd.env_id_offset;

  if (n_total > 0) {
    if (MODE == 1) {
      EnvCtx c = make_ctx(a, env);
      if (a.ext_terminated != nullptr) c.terminated = a.ext_terminated[env] != 0;
      const float v = reward_term(*a.adhoc, S, CS, a, env, c, 0, 64);
      if (valid) a.term_out[env] = v;
      return;
    }
    RL_STAMP(1);                   // index setup done
    if (ph & 0x8000u) return;      // profiling aid: empty launch
    // ---- this group's tasks, straight-line for a baked spec ---------------------------------------------
    const EnvCtx c = a.in[IF_QUAT].ptr != nullptr ? make_ctx(a, env) : EnvCtx{};   // reset-only launches carry no state
    RL_STAMP(2);                   // context (root state loads + 3 quaternion rotations) done
    if (ph & 0x4000u) { if (valid && c.gate < -1.f) a.termv[env] = c.gate; return; }   // profiling aid: context only
    const bool with_dones = (ph & RL_PHASE_DONES) != 0;
    P::template for_tasks<G>(a, g, [&](const Task& tk, const RlRewardTerm& rt, const RlObsTerm& ot, const bool corrupt) {
      if (tk.kind == TK_REWARD) {
        if (!(ph & RL_PHASE_REWARDS)) return;
        const int k = tk.a;
        const float raw = reward_term(rt, S, CS, a, env, c, tk.lo, tk.hi);
        const bool late = (tk.b != 0) || (tk.hi != 64);   // one half of a split term: finalisation adds the halves
        if (late) {
          if (valid) a.termv[(size_t)(2 * k + tk.b) * a.Ncap + env] = raw;
        } else {
          // RewardManager.compute [IL]: value = func * weight * dt; sums += value; step_reward = value / dt
          const float val = (raw * rt.weight) * S.step_dt;
          if (valid) {
            a.termv[(size_t)(2 * k) * a.Ncap + env] = val;
            STF(IF_SUMS, k, LDF(IF_SUMS, k) + val);
            if (a.in[IF_STEPR].ptr) STF(IF_STEPR, k, val / S.step_dt);
          }
        }
      } else if (tk.kind == TK_OBS) {
        if (!(ph & RL_PHASE_OBS) || a.out.obs[tk.a] == nullptr) return;
        const int eplen_eff = do_reset ? 0 : (ld_eplen(a, env) + (with_dones ? 1 : 0));
        obs_task(a, S, CS, ot, corrupt, rs, tk.a, tk.b, tk.col0, tk.lo, tk.hi, env, valid, c, nullptr, eplen_eff, do_reset, xpose);
      } else if (tk.kind == TK_DONES) {
        if (!with_dones) return;
        const uint32_t fl = eval_dones<P>(a, S, env, c);
        if (valid) {
          if (a.out.done_bits) a.out.done_bits[env] = (uint8_t)(fl & 0xff);
          if (a.out.terminated) a.out.terminated[env] = (uint8_t)((fl >> 8) & 1);
          if (a.out.truncated) a.out.truncated[env] = (uint8_t)((fl >> 9) & 1);
        }
      } else if (tk.kind == TK_COMMAND) {
        if (!(ph & (RL_PHASE_COMMAND | RL_PHASE_OBS | RL_PHASE_RESET))) return;
        const auto& cc = P::command(a);
        float c0 = c.c0, c1 = c.c1, c2 = c.c2;
        float mxy = 0.f, myaw = 0.f, tleft = 0.f, head = 0.f;
        int ishead = 0, isstand = 0;
        const bool cmd_state = (ph & (RL_PHASE_COMMAND | RL_PHASE_RESET)) != 0;
        if (cmd_state) {
          mxy = LDF(IF_MXY, 0); myaw = LDF(IF_MYAW, 0); tleft = LDF(IF_TLEFT, 0); head = LDF(IF_HEAD, 0);
          ishead = ld_u8(a.is_heading, env); isstand = ld_u8(a.is_standing, env);
        }
        if (do_reset) {
          // ---- manager reset [IL]: logging partials first (deterministic: lane tree, warp order, block order) ----
          const int bits = (a.out.done_bits != nullptr) ? (int)a.out.done_bits[env] : 0;
          for (int v = 0; v < K + RL_MAX_DONE_TERMS + 2; ++v) {
            float x = 0.f;
            if (valid) {
              if (v < K) x = LDF(IF_SUMS, v);
              else if (v < K + RL_MAX_DONE_TERMS) x = (float)((bits >> (v - K)) & 1);
              else x = (v == K + RL_MAX_DONE_TERMS) ? mxy : myaw;
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
            if (lane == 0) s_red[warp][v] = x;
          }
          __syncthreads();
          if (tid < K + RL_MAX_DONE_TERMS + 2) {
            float tot = 0.f;
            for (int w = 0; w < WPC; ++w) tot += s_red[w][tid];
            a.log_partials[(size_t)qb * RL_LOG_STRIDE + tid] = tot;
          }
          // RewardManager / ActionManager / CommandTerm .reset, episode_length_buf = 0
          if (valid) {
            for (int k = 0; k < K; ++k) STF(IF_SUMS, k, 0.f);
            for (int i = 0; i < S.n_actions; ++i) { STF(IF_ACT, i, 0.f); STF(IF_PACT, i, 0.f); }
            static_cast<int32_t*>(const_cast<void*>(a.in[IF_EPLEN].ptr))[env * a.in[IF_EPLEN].es] = 0;
          }
          mxy = 0.f; myaw = 0.f;
          float u[RL_NUM_CMD_UNIFORMS];
          cmd_uniforms(a, rs, env, RL_STREAM_RESET_COMMAND, u);
          c0 = u[1] * (cc.lin_vel_x_hi - cc.lin_vel_x_lo) + cc.lin_vel_x_lo;
          c1 = u[2] * (cc.lin_vel_y_hi - cc.lin_vel_y_lo) + cc.lin_vel_y_lo;
          c2 = u[3] * (cc.ang_vel_z_hi - cc.ang_vel_z_lo) + cc.ang_vel_z_lo;
          const float keep = (sqrtf(c0 * c0 + c1 * c1) > cc.small_cmd_threshold) ? 1.f : 0.f;
          c0 *= keep; c1 *= keep;
          tleft = u[0] * (cc.resampling_time_hi - cc.resampling_time_lo) + cc.resampling_time_lo;
          if (cc.heading_command) {
            head = u[4] * (cc.heading_hi - cc.heading_lo) + cc.heading_lo;
            ishead = (u[5] <= cc.rel_heading_envs) ? 1 : 0;
          }
          isstand = (u[6] <= cc.rel_standing_envs) ? 1 : 0;
        }
        if (ph & RL_PHASE_COMMAND) {
          // CommandTerm.compute [IL] + UniformThresholdVelocityCommand (V/mdp/commands.py:43-85; the "pits" branch is
          // identically off for the in-scope terrains, V/mdp/utils.py:27-28); withheld from envs done this step
          bool skip = false;
          if ((ph & RL_PHASE_SKIP_DONE_ENVS) && with_dones) skip = ((eval_dones<P>(a, S, env, c) >> 8) & 3) != 0;
          if (!skip) {
            const float dx = c0 - c.vb.x, dy = c1 - c.vb.y;
            mxy = mxy + sqrtf(dx * dx + dy * dy) / cc.max_command_step;
            myaw = myaw + fabsf(c2 - c.wb.z) / cc.max_command_step;
            tleft = tleft - S.step_dt;
            if (tleft <= 0.f) {
              float u[RL_NUM_CMD_UNIFORMS];
              cmd_uniforms(a, rs, env, RL_STREAM_COMMAND, u);
              tleft = u[0] * (cc.resampling_time_hi - cc.resampling_time_lo) + cc.resampling_time_lo;
              c0 = u[1] * (cc.lin_vel_x_hi - cc.lin_vel_x_lo) + cc.lin_vel_x_lo;
              c1 = u[2] * (cc.lin_vel_y_hi - cc.lin_vel_y_lo) + cc.lin_vel_y_lo;
              c2 = u[3] * (cc.ang_vel_z_hi - cc.ang_vel_z_lo) + cc.ang_vel_z_lo;
              if (cc.heading_command) {
                head = u[4] * (cc.heading_hi - cc.heading_lo) + cc.heading_lo;
                ishead = (u[5] <= cc.rel_heading_envs) ? 1 : 0;
              }
              isstand = (u[6] <= cc.rel_standing_envs) ? 1 : 0;
              const float keep = (sqrtf(c0 * c0 + c1 * c1) > cc.small_cmd_threshold) ? 1.f : 0.f;
              c0 *= keep; c1 *= keep;
            }
            if (cc.heading_command && ishead) {
              const V3 fwd = quat_apply(c.qw, c.q, V3{1.f, 0.f, 0.f});
              const float heading = rl_atan2f(fwd.y, fwd.x);
              const float err = wrap_to_pi(head - heading);
              c2 = clampf(cc.heading_control_stiffness * err, cc.ang_vel_z_lo, cc.ang_vel_z_hi);
            }
            if (isstand) { c0 = 0.f; c1 = 0.f; c2 = 0.f; }
          }
        }
        if (cmd_state && valid) {
          STF(IF_MXY, 0, mxy); STF(IF_MYAW, 0, myaw); STF(IF_TLEFT, 0, tleft); STF(IF_HEAD, 0, head);
          st_u8(a.is_heading, env, ishead); st_u8(a.is_standing, env, isstand);
          a.cmd_new[(size_t)0 * a.Ncap + env] = c0; a.cmd_new[(size_t)1 * a.Ncap + env] = c1; a.cmd_new[(size_t)2 * a.Ncap + env] = c2;
        }
        if (ph & RL_PHASE_OBS) {
          const float cmd3[3] = {c0, c1, c2};
          P::for_cmd_obs(a, [&](const RlObsTerm& ct, int og, int oti, int ocol0, bool ocorrupt) {
            if (a.out.obs[og] == nullptr) return;
            obs_task(a, S, CS, ct, ocorrupt, rs, og, oti, ocol0, 0, ct.dim, env, valid, c, cmd3, 0, do_reset, xpose);
          });
        }
      }
    });
  }
  RL_STAMP(3);

  // ---- finalisation of the tile block by the last of its G CTAs ---------------------------------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(a.block_ticket + qb, 1u);
    s_last = (prev == (unsigned)(G - 1));
    if (s_last) a.block_ticket[qb] = 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  RL_STAMP(4);
  uint32_t done_flag = 0;
  if (valid && n_total > 0) {
    if (ph & RL_PHASE_DONES) {
      const int t8 = a.out.terminated ? (int)__ldcg(a.out.terminated + env) : 0;
      const int r8 = a.out.truncated ? (int)__ldcg(a.out.truncated + env) : 0;
      done_flag = (uint32_t)((t8 | r8) != 0);
      int32_t* ep = static_cast<int32_t*>(const_cast<void*>(a.in[IF_EPLEN].ptr)) + env * a.in[IF_EPLEN].es;
      *ep = *ep + 1;
    }
    if (ph & RL_PHASE_REWARDS) {
      const bool terminated = (ph & RL_PHASE_DONES) && a.out.terminated ? (__ldcg(a.out.terminated + env) != 0) : false;
      float total = 0.f;
      P::template for_rewards<G>(a, [&](const RlRewardTerm& t, int k, bool split, bool late) {
        if (t.weight == 0.f) { if (a.in[IF_STEPR].ptr) STF(IF_STEPR, k, 0.f); return; }
        float val;
        if (late) {
          float raw;
          if (t.type == RL_REW_IS_TERMINATED) raw = terminated ? 1.f : 0.f;
          else {
            raw = __ldcg(a.termv + (size_t)(2 * k) * a.Ncap + env);
            if (split) raw = raw + __ldcg(a.termv + (size_t)(2 * k + 1) * a.Ncap + env);
          }
          val = (raw * t.weight) * S.step_dt;
          STF(IF_SUMS, k, LDF(IF_SUMS, k) + val);
          if (a.in[IF_STEPR].ptr) STF(IF_STEPR, k, val / S.step_dt);
        } else {
          val = __ldcg(a.termv + (size_t)(2 * k) * a.Ncap + env);
        }
        total += val;   // manager order
      });
      if (a.out.reward) a.out.reward[env] = total;
    }
    if (ph & (RL_PHASE_COMMAND | RL_PHASE_RESET)) {
      STF(IF_CMD, 0, __ldcg(a.cmd_new + (size_t)0 * a.Ncap + env));
      STF(IF_CMD, 1, __ldcg(a.cmd_new + (size_t)1 * a.Ncap + env));
      STF(IF_CMD, 2, __ldcg(a.cmd_new + (size_t)2 * a.Ncap + env));
    }
  }
  const bool compact = (ph & RL_PHASE_COMPACT) && (ph & RL_PHASE_DONES) && !a.has_ids;
  if (compact) {
    const unsigned m = __ballot_sync(0xffffffffu, done_flag);
    if (lane == 0 && tile < n_tiles) a.tile_mask[tile] = m;
  }
  if (!(compact || do_reset)) return;
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(a.final_ticket, 1u);
    s_final = (prev == (unsigned)(max(n_blocks_live, 1) - 1));
    if (s_final) *a.final_ticket = 0u;
  }
  __syncthreads();
  if (!s_final) return;
  __threadfence();
  RL_STAMP(5);
  if (do_reset && n_total > 0) {
    if (tid < K + RL_MAX_DONE_TERMS + 2) {
      float tot = 0.f;
      for (int b = 0; b < n_blocks_live; ++b) tot += __ldcg(a.log_partials + (size_t)b * RL_LOG_STRIDE + tid);
      const RlResetLog& lg = a.out.reset_log;
      if (tid < K) { if (lg.episode_sum_mean) lg.episode_sum_mean[tid] = tot / (float)n_total; }
      else if (tid < K + RL_MAX_DONE_TERMS) { if (lg.done_term_count) lg.done_term_count[tid - K] = tot; }
      else if (lg.metric_mean) lg.metric_mean[tid - K - RL_MAX_DONE_TERMS] = tot / (float)n_total;
    }
  }
  if (compact) {
    // ordered compaction of reset ids (ManagerBasedRLEnv.step: reset_buf.nonzero() [IL]): thread i owns a
    // contiguous run of tiles -> ids come out ascending
    __shared__ int s_cnt[WPC * 32];
    constexpr int NT = WPC * 32;
    const int per = (n_tiles + NT - 1) / NT;
    const int t0 = tid * per, t1 = min(n_tiles, t0 + per);
    int cnt = 0;
    for (int t = t0; t < t1; ++t) cnt += __popc(__ldcg(a.tile_mask + t));
    s_cnt[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int i = 0; i < NT; ++i) { const int v = s_cnt[i]; s_cnt[i] = run; run += v; }
      if (a.out.n_reset) *a.out.n_reset = run;
    }
    __syncthreads();
    int pos = s_cnt[tid];
    if (a.out.reset_ids)
      for (int t = t0; t < t1; ++t) {
        unsigned m = __ldcg(a.tile_mask + t);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          a.out.reset_ids[pos++] = t * kE + b;
        }
      }
  }
  RL_STAMP(6);
}

// ---------------------------------------------------------------------------------------------------
// process_action: ActionManager.process_action + JointAction.process_actions [IL]
// ---------------------------------------------------------------------------------------------------
__global__ void process_action_kernel(int N, int slot, RlField new_action, RlField action, RlField prev_action,
                                      RlField target, unsigned long long* step_counter, int use_pdl) {
  if (use_pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (step_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *step_counter += 1ull;
  const RlActionCfg& ac = c_spec[slot].action;
  const int A = ac.n_actions;
  const long long total = (long long)N * A;
  const bool env_major = (new_action.env_stride == 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long env; int col;
    if (env_major) { env = i % N; col = (int)(i / N); } else { col = (int)(i % A); env = i / A; }
    const float nv = static_cast<const float*>(new_action.ptr)[env * new_action.env_stride + col * new_action.comp_stride];
    float* ap = static_cast<float*>(action.ptr) + env * action.env_stride + col * action.comp_stride;
    if (prev_action.ptr)
      static_cast<float*>(prev_action.ptr)[env * prev_action.env_stride + col * prev_action.comp_stride] = *ap;
    *ap = nv;
    if (target.ptr) {
      float v = nv * ac.scale[col] + ac.offset[col];
      if (ac.has_clip) v = clampf(v, ac.clip_lo[col], ac.clip_hi[col]);
      static_cast<float*>(target.ptr)[env * target.env_stride + (long long)ac.joint_ids[col] * target.comp_stride] = v;
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
}  // namespace

struct RlCtx {
  int device;
  int slot;
  RlStepSpec spec;
  int G, WPC;             // task groups, warps (= 32-env tiles) per CTA
  Schedule* sched_dev;
  // scratch, grown on demand (outside stream capture)
  int Ncap;
  float* termv;
  float* cmd_new;
  unsigned int* block_ticket;
  unsigned int* final_ticket;
  uint32_t* tile_mask;
  float* log_partials;
  RlRewardTerm* adhoc_dev;
  int sm_count;
  int use_pdl;
  long long* dbg;
  int baked;   // index into RL_BAKED_LIST when the spec equals a build-time specialised one, else -1
};

namespace {

bool to_fd(const RlField& f, FieldD* out) {
  if (f.env_stride > 0x7fffffffLL || f.comp_stride > 0x7fffffffLL || f.env_stride < 0 || f.comp_stride < 0) return false;
  out->ptr = f.ptr; out->es = (int)f.env_stride; out->cs = (int)f.comp_stride;
  return true;
}

// Field descriptors of one launch. `st` may be NULL (reset-only launches).
int fill_args(RlCtx* ctx, KArgs& a, int64_t num_envs, const RlStateView* st, const RlMdpState* mdp, const RlStepOut* out,
              const RlRandom* rnd, uint32_t ph) {
  memset(&a, 0, sizeof(a));
  a.N = (int)num_envs; a.slot = ctx->slot; a.phases = ph;
  a.groups = ctx->G; a.tiles_per_cta = ctx->WPC;
  const int n_tiles = (int)((num_envs + kE - 1) / kE);
  a.n_blocks = (n_tiles + ctx->WPC - 1) / ctx->WPC;
  a.sched = ctx->sched_dev;
  a.termv = ctx->termv; a.cmd_new = ctx->cmd_new; a.Ncap = ctx->Ncap;
  a.block_ticket = ctx->block_ticket; a.final_ticket = ctx->final_ticket; a.tile_mask = ctx->tile_mask;
  a.log_partials = ctx->log_partials; a.use_pdl = ctx->use_pdl; a.dbg = ctx->dbg;
  if (out) a.out = *out;
  if (rnd) a.rnd = *rnd;
  bool ok = true;
  RlField none{nullptr, 0, 0};
  auto S = [&](const RlField RlStateView::*m) -> const RlField& { return st ? st->*m : none; };
  ok &= to_fd(S(&RlStateView::root_pos_w), &a.in[IF_ROOT_POS]); ok &= to_fd(S(&RlStateView::root_quat_w), &a.in[IF_QUAT]);
  ok &= to_fd(S(&RlStateView::root_lin_vel_w), &a.in[IF_LIN_VEL]); ok &= to_fd(S(&RlStateView::root_ang_vel_w), &a.in[IF_ANG_VEL]);
  ok &= to_fd(S(&RlStateView::joint_pos), &a.in[IF_JPOS]); ok &= to_fd(S(&RlStateView::joint_vel), &a.in[IF_JVEL]);
  ok &= to_fd(S(&RlStateView::joint_acc), &a.in[IF_JACC]); ok &= to_fd(S(&RlStateView::applied_torque), &a.in[IF_JTAU]);
  ok &= to_fd(S(&RlStateView::current_air_time), &a.in[IF_CAIR]); ok &= to_fd(S(&RlStateView::last_air_time), &a.in[IF_LAIR]);
  ok &= to_fd(S(&RlStateView::current_contact_time), &a.in[IF_CCON]); ok &= to_fd(S(&RlStateView::last_contact_time), &a.in[IF_LCON]);
  ok &= to_fd(S(&RlStateView::body_pos_w), &a.in[IF_BPOS]); ok &= to_fd(S(&RlStateView::body_lin_vel_w), &a.in[IF_BVEL]);
  ok &= to_fd(S(&RlStateView::ray_sensor_pos_z), &a.in[IF_RAYPOS]);
  ok &= to_fd(S(&RlStateView::net_forces_w_history), &a.in[IF_HIST]); ok &= to_fd(S(&RlStateView::ray_hits_z), &a.in[IF_RAYS]);
  ok &= to_fd(mdp->command, &a.in[IF_CMD]); ok &= to_fd(mdp->heading_target, &a.in[IF_HEAD]);
  ok &= to_fd(mdp->time_left, &a.in[IF_TLEFT]); ok &= to_fd(mdp->metric_error_vel_xy, &a.in[IF_MXY]);
  ok &= to_fd(mdp->metric_error_vel_yaw, &a.in[IF_MYAW]); ok &= to_fd(mdp->episode_length, &a.in[IF_EPLEN]);
  ok &= to_fd(mdp->episode_sums, &a.in[IF_SUMS]);
  ok &= to_fd(mdp->action, &a.in[IF_ACT]); ok &= to_fd(mdp->prev_action, &a.in[IF_PACT]);
  ok &= to_fd(mdp->is_heading_env, &a.is_heading); ok &= to_fd(mdp->is_standing_env, &a.is_standing);
  ok &= to_fd(a.out.step_reward, &a.in[IF_STEPR]);
  if (!ok) return fail(RL_EINVAL, "field strides must be non-negative and below 2^31 elements%s", "");
  return RL_OK;
}

int validate_spec(const RlStepSpec* s) {
  if (s->abi_version != RL_ABI_VERSION) return fail(RL_EINVAL, "spec.abi_version %s%lld != library %lld", "", s->abi_version, RL_ABI_VERSION);
  if (s->num_joints < 1 || s->num_joints > RL_MAX_JOINTS) return fail(RL_EINVAL, "num_joints out of range%s (%lld)", "", s->num_joints);
  if (s->num_hist_bodies < 0 || s->num_hist_bodies > RL_MAX_BODIES) return fail(RL_EINVAL, "num_hist_bodies out of range%s (%lld)", "", s->num_hist_bodies);
  if (s->hist_len < 0 || s->hist_len > 8) return fail(RL_EINVAL, "hist_len out of range%s (%lld)", "", s->hist_len);
  if (s->num_time_bodies < 0 || s->num_time_bodies > RL_MAX_TIME_BODIES) return fail(RL_EINVAL, "num_time_bodies out of range%s (%lld)", "", s->num_time_bodies);
  if (s->num_asset_bodies < 0 || s->num_asset_bodies > RL_MAX_ASSET_BODIES) return fail(RL_EINVAL, "num_asset_bodies out of range%s (%lld)", "", s->num_asset_bodies);
  if (s->num_rays < 0 || s->num_rays > 4096) return fail(RL_EINVAL, "num_rays out of range%s (%lld)", "", s->num_rays);
  if (s->num_reward_terms < 0 || s->num_reward_terms > RL_MAX_REWARD_TERMS) return fail(RL_EINVAL, "num_reward_terms out of range%s (%lld)", "", s->num_reward_terms);
  if (s->num_done_terms < 0 || s->num_done_terms > RL_MAX_DONE_TERMS) return fail(RL_EINVAL, "num_done_terms out of range%s (%lld)", "", s->num_done_terms);
  if (s->action.n_actions < 0 || s->action.n_actions > RL_MAX_JOINTS) return fail(RL_EINVAL, "n_actions out of range%s (%lld)", "", s->action.n_actions);
  if (!(s->step_dt > 0.f)) return fail(RL_EINVAL, "step_dt must be positive%s", "");
  for (int a = 0; a < s->action.n_actions; ++a)
    if (s->action.joint_ids[a] >= s->num_joints) return fail(RL_EINVAL, "action.joint_ids[%s%lld] out of range", "", a);
  for (int k = 0; k < s->num_reward_terms; ++k) {
    const RlRewardTerm& t = s->rewards[k];
    if (t.type <= RL_REW_NONE || t.type >= RL_REW_TYPE_COUNT) return fail(RL_EINVAL, "reward term %s%lld: unknown type %lld", "", k, t.type);
    if (t.n_idx < 0 || t.n_idx > RL_MAX_IDX) return fail(RL_EINVAL, "reward term %s%lld: n_idx out of range", "", k);
    if (t.type == RL_REW_FEET_AIR_TIME_VARIANCE && t.n_idx < 2) return fail(RL_EINVAL, "reward term %s%lld: variance needs >= 2 feet", "", k);
    if (t.type == RL_REW_FEET_GAIT && t.n_idx != 4) return fail(RL_EINVAL, "reward term %s%lld: feet_gait needs two synced pairs", "", k);
  }
  int n_tasks = 2;
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
    const RlObsGroup& G = s->obs[g];
    if (G.n_terms < 0 || G.n_terms > RL_MAX_OBS_TERMS) return fail(RL_EINVAL, "obs group %s%lld: n_terms out of range", "", g);
    int dim = 0;
    for (int t = 0; t < G.n_terms; ++t) {
      const RlObsTerm& o = G.terms[t];
      if (o.type <= RL_OBS_NONE || o.type >= RL_OBS_TYPE_COUNT) return fail(RL_EINVAL, "obs group %s%lld term %lld: unknown type", "", g, t);
      if (o.type == RL_OBS_HEIGHT_SCAN && o.dim != s->num_rays) return fail(RL_EINVAL, "obs group %s%lld: height_scan dim != num_rays", "", g);
      if ((o.type == RL_OBS_JOINT_POS_REL || o.type == RL_OBS_JOINT_VEL_REL || o.type == RL_OBS_JOINT_POS_REL_WITHOUT_WHEEL) && o.dim > RL_MAX_JOINTS)
        return fail(RL_EINVAL, "obs group %s%lld: joint term wider than RL_MAX_JOINTS", "", g);
      dim += o.dim;
      n_tasks += (o.dim + 63) / 64;
    }
    if (dim != G.dim) return fail(RL_EINVAL, "obs group %s%lld: dim %lld != sum of term dims", "", g, G.dim);
  }
  n_tasks += 2 * s->num_reward_terms;
  if (n_tasks > RL_MAX_TASKS) return fail(RL_EUNSUPPORTED, "spec needs %s%lld tasks, the schedule holds %lld", "", n_tasks, RL_MAX_TASKS);
  return RL_OK;
}

template <class P, int G, int WPC, int MODE>
int launch_step(RlCtx* ctx, const KArgs& a, cudaStream_t st) {
  const int grid = MODE == 1 ? a.n_blocks : a.n_blocks * G;
  if (grid <= 0) return RL_OK;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  constexpr size_t smem = (size_t)WPC * 32 * 33 * sizeof(float);
  static thread_local int configured_device = -1;
  if (configured_device != ctx->device) {
    CUDA_TRY(cudaFuncSetAttribute(mdp_step_kernel<P, G, WPC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_device = ctx->device;
  }
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(WPC * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = a.use_pdl ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, mdp_step_kernel<P, G, WPC, MODE>, a));
  return RL_OK;
}

// (task groups, warps per CTA) pairs compiled for the generic kernel / for every baked spec
#define RL_DYN_CONFIGS(X) X(8, 8) X(16, 16)
#define RL_STATIC_CONFIGS(X) X(16, 16) X(24, 16) X(37, 32)

template <class P>
int dispatch_config(RlCtx* ctx, const KArgs& a, cudaStream_t st, bool* found) {
  *found = true;
  if constexpr (P::kStatic) {
#define RL_CASE(G_, W_) if (ctx->G == (G_) && ctx->WPC == (W_)) return launch_step<P, G_, W_, 0>(ctx, a, st);
    RL_STATIC_CONFIGS(RL_CASE)
#undef RL_CASE
  } else {
#define RL_CASE(G_, W_) if (ctx->G == (G_) && ctx->WPC == (W_)) return launch_step<P, G_, W_, 0>(ctx, a, st);
    RL_DYN_CONFIGS(RL_CASE)
#undef RL_CASE
  }
  *found = false;
  return RL_OK;
}

int dispatch_step(RlCtx* ctx, const KArgs& a, cudaStream_t st) {
  bool found = false;
  if (ctx->baked >= 0) {
    int idx = 0, rc = RL_OK;
#define RL_TRY_BAKED(B) if (idx++ == ctx->baked) rc = dispatch_config<StaticPolicy<baked::B>>(ctx, a, st, &found);
    RL_BAKED_LIST(RL_TRY_BAKED)
#undef RL_TRY_BAKED
    if (found) return rc;
  }
  int rc = dispatch_config<DynPolicy>(ctx, a, st, &found);
  if (!found) return fail(RL_EINVAL, "launch config (%s%lld groups, %lld warps per CTA) is not compiled for this spec", "", ctx->G, ctx->WPC);
  return rc;
}

bool config_compiled(const RlCtx* ctx, int G, int W) {
  bool ok = false;
  if (ctx->baked >= 0) {
#define RL_CASE(G_, W_) ok = ok || (G == (G_) && W == (W_));
    RL_STATIC_CONFIGS(RL_CASE)
#undef RL_CASE
  }
#define RL_CASE(G_, W_) ok = ok || (G == (G_) && W == (W_));
  RL_DYN_CONFIGS(RL_CASE)
#undef RL_CASE
  return ok;
}

// scratch sized for num_envs (first call at a larger size must happen outside stream capture)
int ensure_scratch(RlCtx* ctx, int64_t num_envs, cudaStream_t st) {
  if (num_envs <= ctx->Ncap) return RL_OK;
  cudaStreamCaptureStatus cs;
  CUDA_TRY(cudaStreamIsCapturing(st, &cs));
  if (cs != cudaStreamCaptureStatusNone)
    return fail(RL_EINVAL, "the first call at this num_envs must happen outside stream capture (scratch allocation)%s", "");
  CUDA_TRY(cudaDeviceSynchronize());
  if (ctx->termv) cudaFree(ctx->termv);
  if (ctx->cmd_new) cudaFree(ctx->cmd_new);
  if (ctx->block_ticket) cudaFree(ctx->block_ticket);
  if (ctx->tile_mask) cudaFree(ctx->tile_mask);
  if (ctx->log_partials) cudaFree(ctx->log_partials);
  ctx->termv = nullptr; ctx->cmd_new = nullptr; ctx->block_ticket = nullptr; ctx->tile_mask = nullptr; ctx->log_partials = nullptr;
  const int64_t cap = ((num_envs + 1023) / 1024) * 1024;
  const int64_t tiles = cap / kE, blocks = tiles;  // enough for any warps-per-CTA >= 1
  const int K2 = 2 * (ctx->spec.num_reward_terms > 0 ? ctx->spec.num_reward_terms : 1);
  CUDA_TRY(cudaMalloc(&ctx->termv, sizeof(float) * (size_t)K2 * cap));
  CUDA_TRY(cudaMemset(ctx->termv, 0, sizeof(float) * (size_t)K2 * cap));
  CUDA_TRY(cudaMalloc(&ctx->cmd_new, sizeof(float) * 3 * cap));
  CUDA_TRY(cudaMemset(ctx->cmd_new, 0, sizeof(float) * 3 * cap));
  CUDA_TRY(cudaMalloc(&ctx->block_ticket, sizeof(unsigned int) * blocks));
  CUDA_TRY(cudaMemset(ctx->block_ticket, 0, sizeof(unsigned int) * blocks));
  CUDA_TRY(cudaMalloc(&ctx->tile_mask, sizeof(uint32_t) * tiles));
  CUDA_TRY(cudaMemset(ctx->tile_mask, 0, sizeof(uint32_t) * tiles));
  CUDA_TRY(cudaMalloc(&ctx->log_partials, sizeof(float) * (size_t)blocks * RL_LOG_STRIDE));
  CUDA_TRY(cudaMemset(ctx->log_partials, 0, sizeof(float) * (size_t)blocks * RL_LOG_STRIDE));
  ctx->Ncap = (int)cap;
  return RL_OK;
}

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(true) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

bool g_slots[16][RL_SPEC_SLOTS];

int upload_schedule(RlCtx* ctx) {
  const Schedule sc = make_schedule(ctx->spec, ctx->G);
  CUDA_TRY(cudaMemcpy(ctx->sched_dev, &sc, sizeof(Schedule), cudaMemcpyHostToDevice));
  return RL_OK;
}

}  // namespace

extern "C" {

int rl_abi_version(void) { return RL_ABI_VERSION; }
const char* rl_last_error(void) { return g_err; }

int64_t rl_struct_sizeof(const char* name) {
  if (!name) return -1;
#define RL_SZ(T) if (strcmp(name, #T) == 0) return (int64_t)sizeof(T)
  RL_SZ(RlRewardTerm); RL_SZ(RlObsTerm); RL_SZ(RlObsGroup); RL_SZ(RlDoneTerm); RL_SZ(RlCommandCfg);
  RL_SZ(RlActionCfg); RL_SZ(RlStepSpec); RL_SZ(RlField); RL_SZ(RlStateView); RL_SZ(RlMdpState);
  RL_SZ(RlStepOut); RL_SZ(RlRandom); RL_SZ(RlResetLog);
#undef RL_SZ
  return -1;
}

int rl_ctx_create(const RlStepSpec* spec, int device, RlCtx** out) {
  if (!spec || !out) return fail(RL_EINVAL, "rl_ctx_create: null argument%s", "");
  *out = nullptr;
  int rc = validate_spec(spec);
  if (rc != RL_OK) return rc;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev || device >= 16) return fail(RL_EINVAL, "rl_ctx_create: bad device%s %lld", "", device);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(RL_ECUDA, "rl_ctx_create: cudaSetDevice failed%s", "");
  int slot = -1;
  for (int i = 0; i < RL_SPEC_SLOTS; ++i) if (!g_slots[device][i]) { slot = i; break; }
  if (slot < 0) return fail(RL_ENOMEM, "rl_ctx_create: all %s%lld constant-memory spec slots of this device are in use", "", RL_SPEC_SLOTS);
  RlCtx* ctx = new (std::nothrow) RlCtx();
  if (!ctx) return fail(RL_ENOMEM, "rl_ctx_create: out of host memory%s", "");
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->slot = slot;
  ctx->spec = *spec;
  ctx->G = 16;
  ctx->WPC = 16;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaMemcpyToSymbol(c_spec, spec, sizeof(RlStepSpec), sizeof(RlStepSpec) * slot));
  CUDA_TRY(cudaMalloc(&ctx->final_ticket, sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(ctx->final_ticket, 0, sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc(&ctx->adhoc_dev, sizeof(RlRewardTerm) * 64));
  CUDA_TRY(cudaMalloc(&ctx->sched_dev, sizeof(Schedule)));
  rc = upload_schedule(ctx);
  if (rc != RL_OK) return rc;
  rc = ensure_scratch(ctx, 4096, 0);
  if (rc != RL_OK) return rc;
  ctx->baked = -1;
  {
    int idx = 0;
#define RL_MATCH_BAKED(B) if (ctx->baked < 0 && memcmp(spec, &baked::B::spec, sizeof(RlStepSpec)) == 0) ctx->baked = idx; ++idx;
    RL_BAKED_LIST(RL_MATCH_BAKED)
#undef RL_MATCH_BAKED
  }
  g_slots[device][slot] = true;
  *out = ctx;
  return RL_OK;
}

void rl_ctx_destroy(RlCtx* ctx) {
  if (!ctx) return;
  DeviceGuard guard(ctx->device);
  cudaFree(ctx->termv); cudaFree(ctx->cmd_new); cudaFree(ctx->block_ticket); cudaFree(ctx->final_ticket);
  cudaFree(ctx->tile_mask); cudaFree(ctx->log_partials); cudaFree(ctx->adhoc_dev); cudaFree(ctx->sched_dev);
  g_slots[ctx->device][ctx->slot] = false;
  delete ctx;
}

int rl_ctx_set_launch_config(RlCtx* ctx, int groups, int warps_per_cta) {
  if (!ctx) return fail(RL_EINVAL, "null ctx%s", "");
  const int G = groups > 0 ? groups : 16, W = warps_per_cta > 0 ? warps_per_cta : 16;
  if (!config_compiled(ctx, G, W))
    return fail(RL_EINVAL, "launch config (%s%lld groups, %lld warps per CTA) is not compiled (see RL_STATIC_CONFIGS / RL_DYN_CONFIGS)", "", G, W);
  DeviceGuard guard(ctx->device);
  ctx->G = G; ctx->WPC = W;
  return upload_schedule(ctx);  // synchronous: not for hot loops
}

int rl_ctx_set_debug_buffer(RlCtx* ctx, void* device_i64_buffer) {
  if (!ctx) return fail(RL_EINVAL, "null ctx%s", "");
  ctx->dbg = static_cast<long long*>(device_i64_buffer);
  return RL_OK;
}

int rl_ctx_set_pdl(RlCtx* ctx, int enabled) {
  if (!ctx) return fail(RL_EINVAL, "null ctx%s", "");
  ctx->use_pdl = enabled ? 1 : 0;
  return RL_OK;
}

int rl_process_action(RlCtx* ctx, int64_t num_envs, const RlField* new_action, const RlMdpState* mdp,
                      const RlField* joint_target, uint64_t* step_counter, void* stream) {
  if (!ctx || !new_action || !mdp) return fail(RL_EINVAL, "rl_process_action: null argument%s", "");
  if (num_envs <= 0) return RL_OK;
  if (!new_action->ptr || !mdp->action.ptr) return fail(RL_EINVAL, "rl_process_action: action pointers required%s", "");
  DeviceGuard guard(ctx->device);
  RlField tgt = joint_target ? *joint_target : RlField{nullptr, 0, 0};
  const long long total = num_envs * ctx->spec.action.n_actions;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads);
  if (blocks <= 0) return RL_OK;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = ctx->use_pdl ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, process_action_kernel, (int)num_envs, ctx->slot, *new_action, mdp->action,
                              mdp->prev_action, tgt, (unsigned long long*)step_counter, ctx->use_pdl));
  return RL_OK;
}

int rl_step(RlCtx* ctx, int64_t num_envs, const RlStateView* state, const RlMdpState* mdp, const RlStepOut* out,
            const RlRandom* rnd, uint32_t phases, const int32_t* env_ids, const int32_t* n_env_ids, void* stream) {
  if (!ctx || !state || !mdp || !out || !rnd) return fail(RL_EINVAL, "rl_step: null argument%s", "");
  if (num_envs <= 0) return RL_OK;
  if (num_envs > 0x7fffffff / 8) return fail(RL_EINVAL, "rl_step: num_envs too large%s", "");
  if ((env_ids == nullptr) != (n_env_ids == nullptr)) return fail(RL_EINVAL, "rl_step: env_ids and n_env_ids go together%s", "");
  const RlStepSpec& s = ctx->spec;
  if (!state->root_quat_w.ptr || !state->root_lin_vel_w.ptr || !state->root_ang_vel_w.ptr || !state->root_pos_w.ptr ||
      !state->joint_pos.ptr || !state->joint_vel.ptr || !mdp->command.ptr || !mdp->action.ptr || !mdp->episode_length.ptr)
    return fail(RL_EINVAL, "rl_step: root/joint/command/action/episode_length fields are required%s", "");
  if (phases & (RL_PHASE_DONES | RL_PHASE_REWARDS)) {
    if (!state->joint_acc.ptr || !state->applied_torque.ptr || !mdp->prev_action.ptr)
      return fail(RL_EINVAL, "rl_step: joint_acc/applied_torque/prev_action required for rewards%s", "");
    if (s.num_hist_bodies > 0 && !state->net_forces_w_history.ptr) return fail(RL_EINVAL, "rl_step: net_forces_w_history required%s", "");
    if (s.num_time_bodies > 0 && (!state->current_air_time.ptr || !state->last_air_time.ptr || !state->current_contact_time.ptr || !state->last_contact_time.ptr))
      return fail(RL_EINVAL, "rl_step: air/contact time fields required%s", "");
    if (s.num_asset_bodies > 0 && (!state->body_pos_w.ptr || !state->body_lin_vel_w.ptr)) return fail(RL_EINVAL, "rl_step: body_pos_w/body_lin_vel_w required%s", "");
  }
  if ((phases & RL_PHASE_REWARDS) && (!mdp->episode_sums.ptr && s.num_reward_terms > 0))
    return fail(RL_EINVAL, "rl_step: episode_sums required for the reward phase%s", "");
  if (phases & (RL_PHASE_COMMAND | RL_PHASE_RESET)) {
    if (!mdp->heading_target.ptr || !mdp->time_left.ptr || !mdp->is_heading_env.ptr || !mdp->is_standing_env.ptr ||
        !mdp->metric_error_vel_xy.ptr || !mdp->metric_error_vel_yaw.ptr)
      return fail(RL_EINVAL, "rl_step: command state fields required for the command / reset phase%s", "");
  }
  if ((phases & RL_PHASE_OBS) && s.num_rays > 0 && (!state->ray_hits_z.ptr || !state->ray_sensor_pos_z.ptr)) {
    bool needs = false;
    for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
      if (out->obs[g]) for (int t = 0; t < s.obs[g].n_terms; ++t) needs = needs || s.obs[g].terms[t].type == RL_OBS_HEIGHT_SCAN;
    if (needs) return fail(RL_EINVAL, "rl_step: ray_hits_z / ray_sensor_pos_z required for height_scan%s", "");
  }
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
    if (out->obs[g] && out->obs_pitch[g] < s.obs[g].dim) return fail(RL_EINVAL, "rl_step: obs_pitch[%s%lld] smaller than the group dim", "", g);
  if ((phases & RL_PHASE_RESET) && !env_ids) return fail(RL_EINVAL, "rl_step: RL_PHASE_RESET needs env_ids%s", "");
  if ((phases & RL_PHASE_RESET) && (phases & (RL_PHASE_DONES | RL_PHASE_REWARDS))) return fail(RL_EINVAL, "rl_step: RESET combines with COMMAND/OBS only%s", "");
  if ((phases & RL_PHASE_RESET) && (!mdp->episode_sums.ptr || !mdp->prev_action.ptr))
    return fail(RL_EINVAL, "rl_step: RL_PHASE_RESET needs every RlMdpState field%s", "");
  if (env_ids && (phases & RL_PHASE_COMPACT)) return fail(RL_EINVAL, "rl_step: compaction is not available on an env_ids subset%s", "");
  if ((phases & RL_PHASE_COMPACT) && !(phases & RL_PHASE_DONES)) return fail(RL_EINVAL, "rl_step: COMPACT needs DONES%s", "");
  DeviceGuard guard(ctx->device);
  int rc = ensure_scratch(ctx, num_envs, (cudaStream_t)stream);
  if (rc != RL_OK) return rc;
  KArgs a;
  rc = fill_args(ctx, a, num_envs, state, mdp, out, rnd, phases);
  if (rc != RL_OK) return rc;
  a.has_ids = env_ids != nullptr; a.env_ids = env_ids; a.n_env_ids = n_env_ids;
  return dispatch_step(ctx, a, (cudaStream_t)stream);
}

int rl_reset_envs(RlCtx* ctx, int64_t num_envs, const RlMdpState* mdp, const uint8_t* done_bits, const RlRandom* rnd,
                  const RlResetLog* log, const int32_t* env_ids, const int32_t* n_env_ids, void* stream) {
  if (!ctx || !mdp || !rnd || !env_ids || !n_env_ids) return fail(RL_EINVAL, "rl_reset_envs: null argument%s", "");
  if (num_envs <= 0) return RL_OK;
  if (!mdp->episode_sums.ptr || !mdp->action.ptr || !mdp->prev_action.ptr || !mdp->command.ptr || !mdp->time_left.ptr ||
      !mdp->heading_target.ptr || !mdp->is_heading_env.ptr || !mdp->is_standing_env.ptr || !mdp->metric_error_vel_xy.ptr ||
      !mdp->metric_error_vel_yaw.ptr || !mdp->episode_length.ptr)
    return fail(RL_EINVAL, "rl_reset_envs: every RlMdpState field is required%s", "");
  DeviceGuard guard(ctx->device);
  int rc = ensure_scratch(ctx, num_envs, (cudaStream_t)stream);
  if (rc != RL_OK) return rc;
  RlStepOut o;
  memset(&o, 0, sizeof(o));
  o.done_bits = const_cast<uint8_t*>(done_bits);
  if (log) o.reset_log = *log;
  KArgs a;
  rc = fill_args(ctx, a, num_envs, nullptr, mdp, &o, rnd, RL_PHASE_RESET);
  if (rc != RL_OK) return rc;
  a.has_ids = 1; a.env_ids = env_ids; a.n_env_ids = n_env_ids;
  return dispatch_step(ctx, a, (cudaStream_t)stream);
}

int rl_term_eval(RlCtx* ctx, int64_t num_envs, const RlRewardTerm* term, const RlStateView* state, const RlMdpState* mdp,
                 const uint8_t* terminated, float* out, void* stream) {
  if (!ctx || !term || !state || !mdp || !out) return fail(RL_EINVAL, "rl_term_eval: null argument%s", "");
  if (num_envs <= 0) return RL_OK;
  if (term->type <= RL_REW_NONE || term->type >= RL_REW_TYPE_COUNT) return fail(RL_EINVAL, "rl_term_eval: unknown term type%s %lld", "", term->type);
  if (term->n_idx < 0 || term->n_idx > RL_MAX_IDX) return fail(RL_EINVAL, "rl_term_eval: n_idx out of range%s", "");
  if (!state->root_quat_w.ptr || !state->root_lin_vel_w.ptr || !state->root_ang_vel_w.ptr || !state->root_pos_w.ptr ||
      !state->joint_pos.ptr || !state->joint_vel.ptr || !state->joint_acc.ptr || !state->applied_torque.ptr ||
      !mdp->command.ptr || !mdp->action.ptr || !mdp->prev_action.ptr || !mdp->episode_length.ptr)
    return fail(RL_EINVAL, "rl_term_eval: state fields missing%s", "");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  // rotate through a small ring of device copies so that back-to-back async calls do not race
  static thread_local int ring = 0;
  RlRewardTerm* dev = ctx->adhoc_dev + (ring++ & 63);
  CUDA_TRY(cudaMemcpyAsync(dev, term, sizeof(RlRewardTerm), cudaMemcpyHostToDevice, st));
  KArgs a;
  int rc = fill_args(ctx, a, num_envs, state, mdp, nullptr, nullptr, 0);
  if (rc != RL_OK) return rc;
  const int n_tiles = (int)((num_envs + kE - 1) / kE);
  a.n_blocks = (n_tiles + 7) / 8;
  a.adhoc = dev; a.ext_terminated = terminated; a.term_out = out; a.use_pdl = 0;
  return launch_step<DynPolicy, 1, 8, 1>(ctx, a, st);
}

}  // extern "C"
