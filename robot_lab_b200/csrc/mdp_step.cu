// mdp_step.cu - fused per-step MDP pipeline for sm_100a (B200).
//
// One launch evaluates, for every env of a CTA's tile: termination terms, all reward terms (+ episode sums
// and per-term step rewards), the velocity-command update, both observation groups and the ordered
// compaction of reset ids. Reference behaviour restated (paths relative to /root/reference, V/ =
// source/robot_lab/robot_lab/tasks/manager_based/locomotion/velocity/): V/mdp/rewards.py:22-687,
// V/mdp/observations.py:17-35, V/mdp/commands.py:22-85, V/velocity_env_cfg.py:106-254,379-664 and the
// IsaacLab manager loops / upstream terms listed in SURVEY.md Appendix A.
//
// Execution model (HBM-bound, no tensor cores - the arithmetic intensity is ~0.5 FLOP/B):
//   * a CTA owns E consecutive envs; LPE lanes cooperate on one env (joint / body / obs-column loops are
//     strided over the lanes, reductions are xor-shuffles inside the lane group);
//   * load phase: the big AoS sensor rows (contact-force history, height-scan ray hits, noise inputs) are
//     staged into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) while the
//     small per-env fields are gathered with coalesced LDGs into an SoA [word][E] shared-memory record -
//     all global reads of the tile are in flight before any arithmetic starts;
//   * compute phase works on shared memory only and assembles the observation rows there;
//   * store phase: observation rows leave with bulk stores (cp.async.bulk.global.shared::cta), the SoA
//     outputs with coalesced STGs; the last CTA to finish compacts the reset ids from per-CTA bit masks.
//
// Built with -fmad=false on purpose: the reference is eager PyTorch, every op rounds on its own, and not
// contracting a*b+c keeps threshold decisions (contact > 1 N, |cmd| > 0.1, ...) bit-identical.

#include "rl_mdp_step.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>

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
// Shared-memory layout of one CTA tile (word offsets). SoA section: word w of local env e at [w*E + e].
// ---------------------------------------------------------------------------------------------------
struct Layout {
  int root_pos, quat, lin_vel, ang_vel;          // 3, 4, 3, 3
  int jpos, jvel, jacc, jtau;                    // J each
  int act, pact;                                 // A each
  int cmd, head, tleft, ishead, isstand;         // 3, 1, 1, 1, 1
  int mxy, myaw, eplen;                          // 1, 1, 1
  int sums;                                      // K
  int cair, lair, ccon, lcon;                    // Bt each
  int bpos, bvel;                                // Ba*3 each
  int raypos;                                    // 1
  int cmdu;                                      // RL_NUM_CMD_UNIFORMS
  int bmax;                                      // B   scratch: max_t |F_b|
  int rew, flags, stepr;                         // 1, 1, K   outputs
  int soa_words;                                 // total words of the SoA section (per env)
  // AoS rows (word offsets from the start of dynamic smem; each is [E][pitch])
  int hist, hist_pitch;
  int rays, rays_pitch;
  int obs[RL_NUM_OBS_GROUPS], obs_pitch[RL_NUM_OBS_GROUPS];
  int obsu[RL_NUM_OBS_GROUPS];                   // uniforms for noise-as-input (same pitch as obs)
  int total_words;
};

struct KArgs {
  int N;
  int slot;
  uint32_t phases;
  int has_ids;
  RlStateView st;
  RlMdpState mdp;
  RlStepOut out;
  RlRandom rnd;
  const int32_t* env_ids;
  const int32_t* n_env_ids;
  Layout L;
  unsigned int* ticket;
  uint32_t* cta_mask;
  float* log_partials;   // [grid][RL_LOG_STRIDE] per-CTA partial sums of the reset logging reductions
  int use_pdl;
  // single-term evaluation (rl_term_eval)
  const RlRewardTerm* adhoc;
  const uint8_t* ext_terminated;
  float* term_out;
};

// ---------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copies (TMA engine, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

__device__ __forceinline__ uint4 rl_philox(const RlRandom& r, long long env, uint32_t stream, uint32_t block) {
  const unsigned long long genv = (unsigned long long)(env + r.env_id_offset);
  const unsigned long long step = r.step + (r.step_counter ? *r.step_counter : 0ull);
  uint4 ctr = make_uint4((uint32_t)genv, (uint32_t)step, (uint32_t)(step >> 32) ^ (uint32_t)(genv >> 32),
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
__device__ __forceinline__ V3 quat_apply_inverse(float w, V3 q, V3 v) {
  V3 t = cross3(q, v);
  t.x *= 2.f; t.y *= 2.f; t.z *= 2.f;
  V3 c = cross3(q, t);
  return V3{(v.x - w * t.x) + c.x, (v.y - w * t.y) + c.y, (v.z - w * t.z) + c.z};
}
__device__ __forceinline__ V3 quat_apply(float w, V3 q, V3 v) {
  V3 t = cross3(q, v);
  t.x *= 2.f; t.y *= 2.f; t.z *= 2.f;
  V3 c = cross3(q, t);
  return V3{(v.x + w * t.x) + c.x, (v.y + w * t.y) + c.y, (v.z + w * t.z) + c.z};
}
__device__ __forceinline__ float remainder_pos(float a, float b) {  // torch.remainder, b > 0
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

// ---------------------------------------------------------------------------------------------------
// Tile loaders / storers. SoA smem record: sm[off + c*E + e]. Integer fields travel bit-cast in the
// float record.
// ---------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_word(T v) { return __int_as_float((int)v); }
template <> __device__ __forceinline__ float to_word<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T from_word(float w) { return (T)__float_as_int(w); }
template <> __device__ __forceinline__ float from_word<float>(float w) { return w; }

template <int E, typename T>
__device__ __forceinline__ void load_soa(float* sm, int off, const RlField& f, int ncomp, int env0, int nvalid,
                                         const int32_t* ids, int tid, int nthreads) {
  if (f.ptr == nullptr || ncomp <= 0) return;
  const T* __restrict__ p = static_cast<const T*>(f.ptr);
  const int total = ncomp * E;
  const bool env_major = (f.env_stride == 1) || (ncomp == 1);
  for (int i = tid; i < total; i += nthreads) {
    int c, e;
    if (env_major) { e = i % E; c = i / E; } else { c = i % ncomp; e = i / ncomp; }
    if (e < nvalid) {
      const long long env = ids ? (long long)ids[env0 + e] : (long long)(env0 + e);
      sm[off + c * E + e] = to_word<T>(p[env * f.env_stride + (long long)c * f.comp_stride]);
    }
  }
}

template <int E>
__device__ __forceinline__ void load_rows(float* sm, int off, int pitch, const float* ptr, long long es,
                                          long long cs, int ncomp, int env0, int nvalid, const int32_t* ids,
                                          int tid, int nthreads) {
  if (ptr == nullptr || ncomp <= 0) return;
  const int total = ncomp * E;
  const bool env_major = (es == 1);
  for (int i = tid; i < total; i += nthreads) {
    int c, e;
    if (env_major) { e = i % E; c = i / E; } else { c = i % ncomp; e = i / ncomp; }
    if (e < nvalid) {
      const long long env = ids ? (long long)ids[env0 + e] : (long long)(env0 + e);
      sm[off + e * pitch + c] = ptr[env * es + (long long)c * cs];
    }
  }
}

template <int E, typename T>
__device__ __forceinline__ void store_soa(const float* sm, int off, const RlField& f, int ncomp, int env0,
                                          int nvalid, const int32_t* ids, int tid, int nthreads) {
  if (f.ptr == nullptr || ncomp <= 0) return;
  T* __restrict__ p = static_cast<T*>(f.ptr);
  const int total = ncomp * E;
  const bool env_major = (f.env_stride == 1) || (ncomp == 1);
  for (int i = tid; i < total; i += nthreads) {
    int c, e;
    if (env_major) { e = i % E; c = i / E; } else { c = i % ncomp; e = i / ncomp; }
    if (e < nvalid) {
      const long long env = ids ? (long long)ids[env0 + e] : (long long)(env0 + e);
      p[env * f.env_stride + (long long)c * f.comp_stride] = from_word<T>(sm[off + c * E + e]);
    }
  }
}

__device__ __forceinline__ bool bulk_ok(const void* base, long long row_elems, int env0, int E) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(base) + (uintptr_t)env0 * (uintptr_t)row_elems * 4u;
  return ((a & 15u) == 0) && ((((long long)E * row_elems * 4) & 15) == 0);
}

// ---------------------------------------------------------------------------------------------------
// Per-env context shared by all terms
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

template <int LPE>
__device__ __forceinline__ float gsum(float v) {
#pragma unroll
  for (int m = LPE / 2; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
template <int LPE>
__device__ __forceinline__ int gor(int v) {
#pragma unroll
  for (int m = LPE / 2; m > 0; m >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

#define SMF(off, c) sm[(off) + (c) * E + e]

template <int E>
__device__ __forceinline__ bool first_contact(const float* sm, const Layout& L, const RlStepSpec& S, int e, int b) {
  const float t = SMF(L.ccon, b);
  return (t > 0.f) && (t < (S.step_dt + S.contact_time_abs_tol));
}
template <int E>
__device__ __forceinline__ V3 body_vec(const float* sm, int off, int e, int b) {
  return V3{SMF(off, 3 * b + 0), SMF(off, 3 * b + 1), SMF(off, 3 * b + 2)};
}

// One reward term, raw value (no weight, no dt). All lanes of the env group return the same value.
template <int E, int LPE>
__device__ float reward_term(const RlRewardTerm& t, const RlStepSpec& S, const Layout& L, const float* sm,
                             const int e, const int sub, const EnvCtx& c) {
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
      const int off = t.type == RL_REW_JOINT_TORQUES_L2 ? L.jtau : (t.type == RL_REW_JOINT_VEL_L2 ? L.jvel : L.jacc);
      float s = 0.f;
      for (int j = sub; j < J; j += LPE)
        if ((t.joint_mask >> j) & 1ull) { const float v = SMF(off, j); s += v * v; }
      return gsum<LPE>(s);
    }
    case RL_REW_JOINT_DEVIATION_L1: {
      float s = 0.f;
      for (int j = sub; j < J; j += LPE)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jpos, j) - S.default_joint_pos[j]);
      return gsum<LPE>(s);
    }
    case RL_REW_JOINT_POS_LIMITS: {
      float s = 0.f;
      for (int j = sub; j < J; j += LPE)
        if ((t.joint_mask >> j) & 1ull) {
          const float q = SMF(L.jpos, j);
          float o = -fminf(q - S.soft_pos_limit_lo[j], 0.f);
          o += fmaxf(q - S.soft_pos_limit_hi[j], 0.f);
          s += o;
        }
      return gsum<LPE>(s);
    }
    case RL_REW_JOINT_VEL_LIMITS: {
      float s = 0.f;
      for (int j = sub; j < J; j += LPE)
        if ((t.joint_mask >> j) & 1ull)
          s += clampf(fabsf(SMF(L.jvel, j)) - S.soft_vel_limit[j] * t.p[0], 0.f, 1.f);
      return gsum<LPE>(s);
    }
    case RL_REW_JOINT_POWER: {
      float s = 0.f;
      for (int j = sub; j < J; j += LPE)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jvel, j) * SMF(L.jtau, j));
      return gsum<LPE>(s);
    }
    case RL_REW_STAND_STILL: {
      float s = 0.f;
      for (int j = sub; j < J; j += LPE)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jpos, j) - S.default_joint_pos[j]);
      s = gsum<LPE>(s);
      s *= (c.cmd_norm < t.p[0]) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_JOINT_POS_PENALTY: {
      float s = 0.f;
      for (int j = sub; j < J; j += LPE)
        if ((t.joint_mask >> j) & 1ull) { const float d = SMF(L.jpos, j) - S.default_joint_pos[j]; s += d * d; }
      const float running = sqrtf(gsum<LPE>(s));
      const bool moving = (c.cmd_norm > t.p[2]) || (c.vxy_norm > t.p[1]);
      return (moving ? running : t.p[0] * running) * c.gate;
    }
    case RL_REW_JOINT_MIRROR: {
      float s = 0.f;
      for (int i = sub; i < t.n_idx; i += LPE) {
        const float d = SMF(L.jpos, t.idx_a[i]) - SMF(L.jpos, t.idx_b[i]);
        s += d * d;
      }
      s = gsum<LPE>(s);
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_MIRROR: {
      float s = 0.f;
      for (int i = sub; i < t.n_idx; i += LPE) {
        const float d = fabsf(SMF(L.act, t.idx_a[i])) - fabsf(SMF(L.act, t.idx_b[i]));
        s += d * d;
      }
      s = gsum<LPE>(s);
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_SYNC: {
      float r = 0.f;
      for (int g = 0; g < t.n_idx; ++g) {
        const int start = t.idx_b[g], n = t.idx_c[g];
        if (n < 2) continue;
        float m = 0.f;
        for (int i = 0; i < n; ++i) m += fabsf(SMF(L.act, t.idx_a[start + i]));
        m = m / (float)n;
        float v = 0.f;
        for (int i = 0; i < n; ++i) { const float d = fabsf(SMF(L.act, t.idx_a[start + i])) - m; v += d * d; }
        r += v / (float)n;
      }
      return (r * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_RATE_L2: {
      const int A = S.action.n_actions;
      float s = 0.f;
      for (int a = sub; a < A; a += LPE) { const float d = SMF(L.act, a) - SMF(L.pact, a); s += d * d; }
      return gsum<LPE>(s);
    }
    case RL_REW_UNDESIRED_CONTACTS: {
      float s = 0.f;
      for (int b = sub; b < S.num_hist_bodies; b += LPE)
        if (((t.body_mask >> b) & 1ull) && (SMF(L.bmax, b) > t.p[0])) s += 1.f;
      return gsum<LPE>(s) * c.gate;
    }
    case RL_REW_CONTACT_FORCES: {
      float s = 0.f;
      for (int b = sub; b < S.num_hist_bodies; b += LPE)
        if ((t.body_mask >> b) & 1ull) s += fmaxf(SMF(L.bmax, b) - t.p[0], 0.f);
      return gsum<LPE>(s);
    }
    case RL_REW_TRACK_LIN_VEL_XY_EXP: {
      const float dx = c.c0 - c.vb.x, dy = c.c1 - c.vb.y;
      return expf(-(dx * dx + dy * dy) / t.p[0]) * c.gate;
    }
    case RL_REW_TRACK_ANG_VEL_Z_EXP: {
      const float d = c.c2 - c.wb.z;
      return expf(-(d * d) / t.p[0]) * c.gate;
    }
    case RL_REW_TRACK_LIN_VEL_XY_YAW_FRAME_EXP: {
      // yaw_quat [IL] then quat_apply_inverse on the world velocity (V/mdp/rewards.py:60)
      const float yaw = atan2f(2.f * (c.qw * c.q.z + c.q.x * c.q.y), 1.f - 2.f * (c.q.y * c.q.y + c.q.z * c.q.z));
      float yw = cosf(yaw / 2.f), yz = sinf(yaw / 2.f);
      const float nrm = fmaxf(sqrtf(yw * yw + yz * yz), 1e-9f);
      yw = yw / nrm; yz = yz / nrm;
      const V3 v = quat_apply_inverse(yw, V3{0.f, 0.f, yz}, c.vw);
      const float dx = c.c0 - v.x, dy = c.c1 - v.y;
      return expf(-(dx * dx + dy * dy) / t.p[0]) * c.gate;
    }
    case RL_REW_TRACK_ANG_VEL_Z_WORLD_EXP: {
      const float d = c.c2 - c.ww.z;
      return expf(-(d * d) / t.p[0]) * c.gate;
    }
    case RL_REW_FEET_AIR_TIME: {
      float s = 0.f;
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = t.idx_a[i];
        s += (SMF(L.lair, b) - t.p[0]) * (first_contact<E>(sm, L, S, e, b) ? 1.f : 0.f);
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: {
      int n_contact = 0;
      for (int i = 0; i < t.n_idx; ++i) n_contact += (SMF(L.ccon, t.idx_a[i]) > 0.f) ? 1 : 0;
      const bool single = (n_contact == 1);
      float r = INFINITY;
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = t.idx_a[i];
        const float ct = SMF(L.ccon, b);
        const float mode = (ct > 0.f) ? ct : SMF(L.cair, b);
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
        const int off = which == 0 ? L.lair : L.lcon;
        float mean = 0.f, m2 = 0.f;
        for (int i = 0; i < t.n_idx; ++i) {
          const float x = fminf(SMF(off, t.idx_a[i]), 0.5f);
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
      auto sync = [&](int a, int b) {
        const float da = SMF(L.cair, a) - SMF(L.cair, b);
        const float dc = SMF(L.ccon, a) - SMF(L.ccon, b);
        return expf(-(fminf(da * da, me2) + fminf(dc * dc, me2)) / sd);
      };
      auto async = [&](int a, int b) {
        const float d0 = SMF(L.cair, a) - SMF(L.ccon, b);
        const float d1 = SMF(L.ccon, a) - SMF(L.cair, b);
        return expf(-(fminf(d0 * d0, me2) + fminf(d1 * d1, me2)) / sd);
      };
      const float sync_r = sync(f00, f01) * sync(f10, f11);
      const float async_r = ((async(f00, f10) * async(f01, f11)) * async(f00, f11)) * async(f10, f01);
      const bool moving = (c.cmd_norm > t.p[3]) || (c.vxy_norm > t.p[2]);
      return (moving ? sync_r * async_r : 0.f) * c.gate;
    }
    case RL_REW_FEET_CONTACT: {
      int n = 0;
      for (int i = 0; i < t.n_idx; ++i) n += first_contact<E>(sm, L, S, e, t.idx_a[i]) ? 1 : 0;
      float r = ((float)n != t.p[0]) ? 1.f : 0.f;
      r *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_CONTACT_WITHOUT_CMD: {
      int n = 0;
      for (int i = 0; i < t.n_idx; ++i) n += first_contact<E>(sm, L, S, e, t.idx_a[i]) ? 1 : 0;
      float r = (float)n;
      r *= (c.cmd_norm < 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_STUMBLE: {
      bool any = false;
      const float* h = sm + L.hist + e * L.hist_pitch;  // t = 0 is the newest sample = net_forces_w
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = t.idx_c[i];
        const float fx = h[3 * b + 0], fy = h[3 * b + 1], fz = h[3 * b + 2];
        any = any || (sqrtf(fx * fx + fy * fy) > 4.f * fabsf(fz));
      }
      return (any ? 1.f : 0.f) * c.gate;
    }
    case RL_REW_FEET_SLIDE: {
      float s = 0.f;
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 vw = body_vec<E>(sm, L.bvel, e, t.idx_b[i]);
        const V3 vb = quat_apply_inverse(c.qw, c.q, V3{vw.x - c.vw.x, vw.y - c.vw.y, vw.z - c.vw.z});
        const float lat = sqrtf(vb.x * vb.x + vb.y * vb.y);
        s += lat * ((SMF(L.bmax, t.idx_c[i]) > 1.0f) ? 1.f : 0.f);
      }
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT: {
      float s = 0.f;
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 p = body_vec<E>(sm, L.bpos, e, t.idx_b[i]);
        const V3 v = body_vec<E>(sm, L.bvel, e, t.idx_b[i]);
        const float d = p.z - t.p[0];
        s += (d * d) * tanhf(t.p[1] * sqrtf(v.x * v.x + v.y * v.y));
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT_BODY: {
      float s = 0.f;
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec<E>(sm, L.bpos, e, t.idx_b[i]);
        const V3 vw = body_vec<E>(sm, L.bvel, e, t.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const V3 vb = quat_apply_inverse(c.qw, c.q, V3{vw.x - c.vw.x, vw.y - c.vw.y, vw.z - c.vw.z});
        const float d = pb.z - t.p[0];
        s += (d * d) * tanhf(t.p[1] * sqrtf(vb.x * vb.x + vb.y * vb.y));
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_DISTANCE_Y_EXP: {
      float s = 0.f;
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec<E>(sm, L.bpos, e, t.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const float want = (t.p[0] / 2.f) * ((i % 2 == 0) ? 1.f : -1.f);
        const float d = want - pb.y;
        s += d * d;
      }
      return expf(-s / t.p[1]) * c.gate;
    }
    case RL_REW_FEET_DISTANCE_XY_EXP: {
      float s = 0.f;
      for (int i = 0; i < 4; ++i) {
        const V3 pw = body_vec<E>(sm, L.bpos, e, t.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const float wx = (i < 2) ? (t.p[1] / 2.f) : (-t.p[1] / 2.f);
        const float wy = (i % 2 == 0) ? (t.p[0] / 2.f) : (-t.p[0] / 2.f);
        const float dx = wx - pb.x, dy = wy - pb.y;
        s += dx * dx + dy * dy;
      }
      return expf(-s / t.p[2]) * c.gate;
    }
    case RL_REW_WHEEL_VEL_PENALTY: {
      float run = 0.f, stand = 0.f;
      for (int i = 0; i < t.n_idx; ++i) {
        const float jv = fabsf(SMF(L.jvel, t.idx_b[i]));
        const float ta = SMF(L.cair, t.idx_a[i]);
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

template <int E>
__device__ __forceinline__ EnvCtx make_ctx(const float* sm, const Layout& L, int e) {
  EnvCtx c;
  c.qw = SMF(L.quat, 0);
  c.q = V3{SMF(L.quat, 1), SMF(L.quat, 2), SMF(L.quat, 3)};
  c.pos = V3{SMF(L.root_pos, 0), SMF(L.root_pos, 1), SMF(L.root_pos, 2)};
  c.vw = V3{SMF(L.lin_vel, 0), SMF(L.lin_vel, 1), SMF(L.lin_vel, 2)};
  c.ww = V3{SMF(L.ang_vel, 0), SMF(L.ang_vel, 1), SMF(L.ang_vel, 2)};
  c.g = quat_apply_inverse(c.qw, c.q, V3{0.f, 0.f, -1.f});
  c.vb = quat_apply_inverse(c.qw, c.q, c.vw);
  c.wb = quat_apply_inverse(c.qw, c.q, c.ww);
  c.gate = clampf(-c.g.z, 0.f, 0.7f) / 0.7f;
  c.c0 = SMF(L.cmd, 0); c.c1 = SMF(L.cmd, 1); c.c2 = SMF(L.cmd, 2);
  c.cmd_norm = sqrtf((c.c0 * c.c0 + c.c1 * c.c1) + c.c2 * c.c2);
  c.vxy_norm = sqrtf(c.vb.x * c.vb.x + c.vb.y * c.vb.y);
  c.terminated = false;
  return c;
}

// max over the history of |F_b| for the lane group's bodies -> smem scratch (shared by 5 terms)
template <int E, int LPE>
__device__ __forceinline__ void body_max_norm(float* sm, const Layout& L, const RlStepSpec& S, int e, int sub) {
  const int B = S.num_hist_bodies, T = S.hist_len;
  const float* h = sm + L.hist + e * L.hist_pitch;
  for (int b = sub; b < B; b += LPE) {
    float m = 0.f;
    for (int t = 0; t < T; ++t) {
      const float* f = h + (t * B + b) * 3;
      const float n = sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
      m = (t == 0) ? n : fmaxf(m, n);
    }
    SMF(L.bmax, b) = m;
  }
}

// CommandTerm.compute [IL] + UniformThresholdVelocityCommand (V/mdp/commands.py:43-85; the "pits" branch
// is identically off for the in-scope terrains, V/mdp/utils.py:27-28). Writes back into the smem record.
template <int E>
__device__ __forceinline__ void command_update(float* sm, const Layout& L, const RlStepSpec& S, const KArgs& a,
                                               int e, long long env, const EnvCtx& c, bool write) {
  const RlCommandCfg& cc = S.command;
  float c0 = c.c0, c1 = c.c1, c2 = c.c2;
  // metrics use the command and state of this step
  {
    const float dx = c0 - c.vb.x, dy = c1 - c.vb.y;
    const float exy = sqrtf(dx * dx + dy * dy) / cc.max_command_step;
    const float eyaw = fabsf(c2 - c.wb.z) / cc.max_command_step;
    if (write) { SMF(L.mxy, 0) = SMF(L.mxy, 0) + exy; SMF(L.myaw, 0) = SMF(L.myaw, 0) + eyaw; }
  }
  float tleft = SMF(L.tleft, 0) - S.step_dt;
  float head = SMF(L.head, 0);
  int ishead = __float_as_int(SMF(L.ishead, 0));
  int isstand = __float_as_int(SMF(L.isstand, 0));
  if (tleft <= 0.f) {
    float u[RL_NUM_CMD_UNIFORMS];
    if (a.rnd.cmd_uniforms != nullptr) {
#pragma unroll
      for (int i = 0; i < RL_NUM_CMD_UNIFORMS; ++i) u[i] = SMF(L.cmdu, i);
    } else {
      const uint4 r0 = rl_philox(a.rnd, env, RL_STREAM_COMMAND, 0), r1 = rl_philox(a.rnd, env, RL_STREAM_COMMAND, 1);
      u[0] = u01(r0.x); u[1] = u01(r0.y); u[2] = u01(r0.z); u[3] = u01(r0.w);
      u[4] = u01(r1.x); u[5] = u01(r1.y); u[6] = u01(r1.z);
    }
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
    const float heading = atan2f(fwd.y, fwd.x);
    const float err = wrap_to_pi(head - heading);
    c2 = clampf(cc.heading_control_stiffness * err, cc.ang_vel_z_lo, cc.ang_vel_z_hi);
  }
  if (isstand) { c0 = 0.f; c1 = 0.f; c2 = 0.f; }
  if (write) {
    SMF(L.cmd, 0) = c0; SMF(L.cmd, 1) = c1; SMF(L.cmd, 2) = c2;
    SMF(L.tleft, 0) = tleft; SMF(L.head, 0) = head;
    SMF(L.ishead, 0) = __int_as_float(ishead); SMF(L.isstand, 0) = __int_as_float(isstand);
  }
}

// One observation group for one env: ObservationManager.compute_group [IL] (clone, +noise, clip, scale, cat)
template <int E, int LPE>
__device__ __forceinline__ void obs_group(float* sm, const Layout& L, const RlStepSpec& S, const KArgs& a, int g,
                                          int e, int sub, long long env, const EnvCtx& c) {
  const RlObsGroup& G = S.obs[g];
  float* row = sm + L.obs[g] + e * L.obs_pitch[g];
  const float* urow = sm + L.obsu[g] + e * L.obs_pitch[g];
  const bool ext_u = (a.rnd.obs_uniforms[g] != nullptr);
  int col0 = 0;
  for (int ti = 0; ti < G.n_terms; ++ti) {
    const RlObsTerm& t = G.terms[ti];
    const bool noisy = t.has_noise && G.enable_corruption;
    for (int qd = sub; qd * 4 < t.dim; qd += LPE) {
      float u4[4] = {0.f, 0.f, 0.f, 0.f};
      if (noisy && !ext_u) {
        const uint4 r = rl_philox(a.rnd, env, RL_STREAM_OBS + g * RL_MAX_OBS_TERMS + ti, (uint32_t)qd);
        u4[0] = u01(r.x); u4[1] = u01(r.y); u4[2] = u01(r.z); u4[3] = u01(r.w);
      }
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        const int col = qd * 4 + r4;
        if (col >= t.dim) break;
        float v;
        switch (t.type) {
          case RL_OBS_BASE_LIN_VEL: v = col == 0 ? c.vb.x : (col == 1 ? c.vb.y : c.vb.z); break;
          case RL_OBS_BASE_ANG_VEL: v = col == 0 ? c.wb.x : (col == 1 ? c.wb.y : c.wb.z); break;
          case RL_OBS_PROJECTED_GRAVITY: v = col == 0 ? c.g.x : (col == 1 ? c.g.y : c.g.z); break;
          case RL_OBS_GENERATED_COMMANDS: v = SMF(L.cmd, col); break;
          case RL_OBS_JOINT_POS_REL: v = SMF(L.jpos, t.ids[col]) - S.default_joint_pos[t.ids[col]]; break;
          case RL_OBS_JOINT_POS_REL_WITHOUT_WHEEL:
            v = SMF(L.jpos, t.ids[col]) - S.default_joint_pos[t.ids[col]];
            if ((t.zero_mask >> col) & 1ull) v = 0.f;
            break;
          case RL_OBS_JOINT_VEL_REL: v = SMF(L.jvel, t.ids[col]) - S.default_joint_vel[t.ids[col]]; break;
          case RL_OBS_LAST_ACTION: v = SMF(L.act, col); break;
          case RL_OBS_HEIGHT_SCAN:
            v = (SMF(L.raypos, 0) - sm[L.rays + e * L.rays_pitch + col]) - t.p[0];
            break;
          case RL_OBS_PHASE: {
            const float ph = ((float)__float_as_int(SMF(L.eplen, 0)) * S.step_dt) / t.p[0];
            v = col == 0 ? sinf((2.f * RL_PI_F) * ph) : cosf((2.f * RL_PI_F) * ph);
            break;
          }
          default: v = 0.f;
        }
        if (noisy) {
          const float u = ext_u ? urow[col0 + col] : u4[r4];
          v = (v + u * (t.noise_hi - t.noise_lo)) + t.noise_lo;
        }
        if (t.has_clip) v = clampf(v, t.clip_lo, t.clip_hi);
        if (t.has_scale) v = v * t.scale;
        row[col0 + col] = v;
      }
    }
    col0 += t.dim;
  }
}

// ---------------------------------------------------------------------------------------------------
// The fused step kernel. MODE 0 = step, MODE 1 = single-term evaluation.
// ---------------------------------------------------------------------------------------------------
template <int E, int LPE, int MODE>
__global__ void __launch_bounds__(E* LPE) mdp_step_kernel(const KArgs a) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_last;
  const RlStepSpec& S = c_spec[a.slot];
  const Layout& L = a.L;
  constexpr int NT = E * LPE;
  const int tid = threadIdx.x;
  const uint32_t ph = a.phases;
  if (a.use_pdl) {
    // launch-latency overlap only: every read below may depend on the predecessor, so wait first
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int n_total = a.has_ids ? *a.n_env_ids : a.N;
  const int env0 = blockIdx.x * E;
  if (MODE == 0 && (ph & RL_PHASE_RESET) && a.has_ids && n_total == 0 && blockIdx.x == 0 && tid < RL_LOG_STRIDE) {
    // nothing to reset: the logged scalars are defined as 0
    const int K0 = S.num_reward_terms;
    if (tid < K0) { if (a.out.reset_log.episode_sum_mean) a.out.reset_log.episode_sum_mean[tid] = 0.f; }
    else if (tid < K0 + RL_MAX_DONE_TERMS) { if (a.out.reset_log.done_term_count) a.out.reset_log.done_term_count[tid - K0] = 0.f; }
    else if (tid < K0 + RL_MAX_DONE_TERMS + 2) { if (a.out.reset_log.metric_mean) a.out.reset_log.metric_mean[tid - K0 - RL_MAX_DONE_TERMS] = 0.f; }
  }
  if (env0 >= n_total && !(a.phases & RL_PHASE_COMPACT)) return;
  const int nvalid = max(0, min(E, n_total - env0));
  const int32_t* ids = a.has_ids ? a.env_ids : nullptr;
  const int J = S.num_joints, A = S.action.n_actions, K = S.num_reward_terms;
  const int Bt = S.num_time_bodies, Ba = S.num_asset_bodies, R = S.num_rays;
  const int HW = S.hist_len * S.num_hist_bodies * 3;
  const bool need_hist = (MODE == 1) || (ph & (RL_PHASE_DONES | RL_PHASE_REWARDS));
  const bool need_rays = (MODE == 0) && (ph & RL_PHASE_OBS) && R > 0;
  const bool do_reset = (MODE == 0) && (ph & RL_PHASE_RESET) && a.has_ids;

  // ---- load phase --------------------------------------------------------------------------------
  const bool full = (nvalid == E) && (ids == nullptr);
  const bool hist_bulk = need_hist && full && HW > 0 && a.st.net_forces_w_history.comp_stride == 1 &&
                         a.st.net_forces_w_history.env_stride == HW &&
                         bulk_ok(a.st.net_forces_w_history.ptr, HW, env0, E);
  const bool rays_bulk = need_rays && full && a.st.ray_hits_z.comp_stride == 1 && a.st.ray_hits_z.env_stride == R &&
                         bulk_ok(a.st.ray_hits_z.ptr, R, env0, E);
  bool obsu_bulk[RL_NUM_OBS_GROUPS];
#pragma unroll
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
    obsu_bulk[g] = (MODE == 0) && (ph & RL_PHASE_OBS) && full && a.rnd.obs_uniforms[g] != nullptr &&
                   S.obs[g].dim > 0 && bulk_ok(a.rnd.obs_uniforms[g], S.obs[g].dim, env0, E);
  if (nvalid > 0) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      uint32_t bytes = 0;
      if (hist_bulk) bytes += (uint32_t)(E * HW * 4);
      if (rays_bulk) bytes += (uint32_t)(E * R * 4);
#pragma unroll
      for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
        if (obsu_bulk[g]) bytes += (uint32_t)(E * S.obs[g].dim * 4);
      mbar_expect_tx(&s_bar, bytes);
      if (hist_bulk)
        bulk_g2s(sm + L.hist, static_cast<const float*>(a.st.net_forces_w_history.ptr) + (size_t)env0 * HW,
                 (uint32_t)(E * HW * 4), &s_bar);
      if (rays_bulk)
        bulk_g2s(sm + L.rays, static_cast<const float*>(a.st.ray_hits_z.ptr) + (size_t)env0 * R,
                 (uint32_t)(E * R * 4), &s_bar);
#pragma unroll
      for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
        if (obsu_bulk[g])
          bulk_g2s(sm + L.obsu[g], a.rnd.obs_uniforms[g] + (size_t)env0 * S.obs[g].dim,
                   (uint32_t)(E * S.obs[g].dim * 4), &s_bar);
    }
    load_soa<E, float>(sm, L.root_pos, a.st.root_pos_w, 3, env0, nvalid, ids, tid, NT);
    load_soa<E, float>(sm, L.quat, a.st.root_quat_w, 4, env0, nvalid, ids, tid, NT);
    load_soa<E, float>(sm, L.lin_vel, a.st.root_lin_vel_w, 3, env0, nvalid, ids, tid, NT);
    load_soa<E, float>(sm, L.ang_vel, a.st.root_ang_vel_w, 3, env0, nvalid, ids, tid, NT);
    load_soa<E, float>(sm, L.jpos, a.st.joint_pos, J, env0, nvalid, ids, tid, NT);
    load_soa<E, float>(sm, L.jvel, a.st.joint_vel, J, env0, nvalid, ids, tid, NT);
    load_soa<E, float>(sm, L.act, a.mdp.action, A, env0, nvalid, ids, tid, NT);
    load_soa<E, float>(sm, L.cmd, a.mdp.command, 3, env0, nvalid, ids, tid, NT);
    load_soa<E, int32_t>(sm, L.eplen, a.mdp.episode_length, 1, env0, nvalid, ids, tid, NT);
    if (need_hist) {
      load_soa<E, float>(sm, L.jacc, a.st.joint_acc, J, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.jtau, a.st.applied_torque, J, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.pact, a.mdp.prev_action, A, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.cair, a.st.current_air_time, Bt, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.lair, a.st.last_air_time, Bt, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.ccon, a.st.current_contact_time, Bt, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.lcon, a.st.last_contact_time, Bt, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.bpos, a.st.body_pos_w, Ba * 3, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.bvel, a.st.body_lin_vel_w, Ba * 3, env0, nvalid, ids, tid, NT);
      if (MODE == 0) load_soa<E, float>(sm, L.sums, a.mdp.episode_sums, K, env0, nvalid, ids, tid, NT);
      if (!hist_bulk)
        load_rows<E>(sm, L.hist, L.hist_pitch, static_cast<const float*>(a.st.net_forces_w_history.ptr),
                     a.st.net_forces_w_history.env_stride, a.st.net_forces_w_history.comp_stride, HW, env0, nvalid,
                     ids, tid, NT);
    }
    if (MODE == 0 && (ph & RL_PHASE_COMMAND)) {
      load_soa<E, float>(sm, L.head, a.mdp.heading_target, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.tleft, a.mdp.time_left, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, uint8_t>(sm, L.ishead, a.mdp.is_heading_env, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, uint8_t>(sm, L.isstand, a.mdp.is_standing_env, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.mxy, a.mdp.metric_error_vel_xy, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.myaw, a.mdp.metric_error_vel_yaw, 1, env0, nvalid, ids, tid, NT);
      if (a.rnd.cmd_uniforms != nullptr) {
        RlField f{const_cast<float*>(a.rnd.cmd_uniforms), 1, (int64_t)a.N};
        load_soa<E, float>(sm, L.cmdu, f, RL_NUM_CMD_UNIFORMS, env0, nvalid, ids, tid, NT);
      }
    }
    if (do_reset) {
      load_soa<E, float>(sm, L.sums, a.mdp.episode_sums, K, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.mxy, a.mdp.metric_error_vel_xy, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.myaw, a.mdp.metric_error_vel_yaw, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, float>(sm, L.head, a.mdp.heading_target, 1, env0, nvalid, ids, tid, NT);
      load_soa<E, uint8_t>(sm, L.ishead, a.mdp.is_heading_env, 1, env0, nvalid, ids, tid, NT);
      if (a.out.done_bits != nullptr) {
        RlField f{a.out.done_bits, 1, 0};
        load_soa<E, uint8_t>(sm, L.flags, f, 1, env0, nvalid, ids, tid, NT);
      }
      if (a.rnd.cmd_uniforms != nullptr && !(ph & RL_PHASE_COMMAND)) {
        RlField f{const_cast<float*>(a.rnd.cmd_uniforms), 1, (int64_t)a.N};
        load_soa<E, float>(sm, L.cmdu, f, RL_NUM_CMD_UNIFORMS, env0, nvalid, ids, tid, NT);
      }
    }
    if (need_rays) {
      load_soa<E, float>(sm, L.raypos, a.st.ray_sensor_pos_z, 1, env0, nvalid, ids, tid, NT);
      if (!rays_bulk)
        load_rows<E>(sm, L.rays, L.rays_pitch, static_cast<const float*>(a.st.ray_hits_z.ptr),
                     a.st.ray_hits_z.env_stride, a.st.ray_hits_z.comp_stride, R, env0, nvalid, ids, tid, NT);
    }
    if (MODE == 0 && (ph & RL_PHASE_OBS)) {
#pragma unroll
      for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
        if (a.rnd.obs_uniforms[g] != nullptr && !obsu_bulk[g])
          load_rows<E>(sm, L.obsu[g], L.obs_pitch[g], a.rnd.obs_uniforms[g], S.obs[g].dim, 1, S.obs[g].dim, env0,
                       nvalid, ids, tid, NT);
    }
    __syncthreads();            // smem record + mbarrier init visible
    mbar_wait(&s_bar, 0);       // bulk copies landed
  }

  // ---- compute phase ----------------------------------------------------------------------------
  const int e = tid / LPE;
  const int sub = tid % LPE;
  const bool valid = e < nvalid;
  const long long env = valid ? (ids ? (long long)ids[env0 + e] : (long long)(env0 + e)) : 0;
  uint32_t done_any = 0;
  if (nvalid > 0) {
    if (do_reset) {
      // -- logging partials of this CTA (summed in CTA order by the last CTA -> deterministic) --
      if (tid < K + RL_MAX_DONE_TERMS + 2) {
        float acc = 0.f;
        for (int el = 0; el < nvalid; ++el) {
          if (tid < K) acc += sm[L.sums + tid * E + el];
          else if (tid < K + RL_MAX_DONE_TERMS)
            acc += (a.out.done_bits != nullptr) ? (float)((__float_as_int(sm[L.flags + el]) >> (tid - K)) & 1) : 0.f;
          else acc += sm[(tid == K + RL_MAX_DONE_TERMS ? L.mxy : L.myaw) + el];
        }
        a.log_partials[(size_t)blockIdx.x * RL_LOG_STRIDE + tid] = acc;
      }
      __syncthreads();
      // -- manager resets: RewardManager / ActionManager / CommandTerm .reset [IL], episode_length_buf = 0 --
      for (int i = tid; i < E * K; i += NT) sm[L.sums + i] = 0.f;
      for (int i = tid; i < E * A; i += NT) { sm[L.act + i] = 0.f; sm[L.pact + i] = 0.f; }
      if (tid < E) {
        const int el = tid;
        sm[L.mxy + el] = 0.f; sm[L.myaw + el] = 0.f; sm[L.eplen + el] = __int_as_float(0);
        const RlCommandCfg& cc = S.command;
        const long long ev = ids ? (long long)ids[env0 + min(el, nvalid - 1)] : (long long)(env0 + el);
        float u[RL_NUM_CMD_UNIFORMS];
        if (a.rnd.cmd_uniforms != nullptr) {
#pragma unroll
          for (int q = 0; q < RL_NUM_CMD_UNIFORMS; ++q) u[q] = sm[L.cmdu + q * E + el];
        } else {
          const uint4 r0 = rl_philox(a.rnd, ev, RL_STREAM_RESET_COMMAND, 0), r1 = rl_philox(a.rnd, ev, RL_STREAM_RESET_COMMAND, 1);
          u[0] = u01(r0.x); u[1] = u01(r0.y); u[2] = u01(r0.z); u[3] = u01(r0.w);
          u[4] = u01(r1.x); u[5] = u01(r1.y); u[6] = u01(r1.z);
        }
        float c0 = u[1] * (cc.lin_vel_x_hi - cc.lin_vel_x_lo) + cc.lin_vel_x_lo;
        float c1 = u[2] * (cc.lin_vel_y_hi - cc.lin_vel_y_lo) + cc.lin_vel_y_lo;
        const float c2 = u[3] * (cc.ang_vel_z_hi - cc.ang_vel_z_lo) + cc.ang_vel_z_lo;
        const float keep = (sqrtf(c0 * c0 + c1 * c1) > cc.small_cmd_threshold) ? 1.f : 0.f;
        c0 *= keep; c1 *= keep;
        sm[L.cmd + 0 * E + el] = c0; sm[L.cmd + 1 * E + el] = c1; sm[L.cmd + 2 * E + el] = c2;
        sm[L.tleft + el] = u[0] * (cc.resampling_time_hi - cc.resampling_time_lo) + cc.resampling_time_lo;
        if (cc.heading_command) {
          sm[L.head + el] = u[4] * (cc.heading_hi - cc.heading_lo) + cc.heading_lo;
          sm[L.ishead + el] = __int_as_float((u[5] <= cc.rel_heading_envs) ? 1 : 0);
        }
        sm[L.isstand + el] = __int_as_float((u[6] <= cc.rel_standing_envs) ? 1 : 0);
      }
      __syncthreads();
    }
    EnvCtx c = make_ctx<E>(sm, L, e);
    if (need_hist) {
      body_max_norm<E, LPE>(sm, L, S, e, sub);
      __syncwarp();
    }
    if (MODE == 1) {
      if (a.ext_terminated != nullptr && valid) c.terminated = a.ext_terminated[env] != 0;
      const float v = reward_term<E, LPE>(*a.adhoc, S, L, sm, e, sub, c);
      if (valid && sub == 0) a.term_out[env] = v;
      return;
    }
    uint32_t bits = 0, term = 0, trunc = 0;
    if (ph & RL_PHASE_DONES) {
      const int eplen = __float_as_int(SMF(L.eplen, 0)) + 1;
      for (int d = 0; d < S.num_done_terms; ++d) {
        const RlDoneTerm& t = S.dones[d];
        int fired = 0;
        if (t.type == RL_DONE_TIME_OUT) {
          fired = eplen >= S.max_episode_length;
        } else if (t.type == RL_DONE_TERRAIN_OUT_OF_BOUNDS) {
          fired = (t.p[2] != 0.f) && ((fabsf(c.pos.x) > t.p[0]) || (fabsf(c.pos.y) > t.p[1]));
        } else if (t.type == RL_DONE_ILLEGAL_CONTACT) {
          int hit = 0;
          for (int b = sub; b < S.num_hist_bodies; b += LPE)
            if (((t.body_mask >> b) & 1ull) && (SMF(L.bmax, b) > t.p[0])) hit = 1;
          fired = gor<LPE>(hit);
        }
        if (fired) { bits |= 1u << d; if (t.time_out) trunc = 1; else term = 1; }
      }
      c.terminated = term != 0;
      __syncwarp();
      if (sub == 0) {
        SMF(L.eplen, 0) = __int_as_float(eplen);
        SMF(L.flags, 0) = __int_as_float((int)(bits | (term << 8) | (trunc << 9)));
      }
    }
    done_any = term | trunc;
    if (ph & RL_PHASE_REWARDS) {
      float total = 0.f;
      for (int k = 0; k < K; ++k) {
        const RlRewardTerm& t = S.rewards[k];
        float val = 0.f, per_dt = 0.f;
        if (t.weight != 0.f) {
          const float raw = reward_term<E, LPE>(t, S, L, sm, e, sub, c);
          val = (raw * t.weight) * S.step_dt;  // RewardManager.compute [IL]: func * weight * dt
          per_dt = val / S.step_dt;
          total += val;
        }
        if (sub == (k % LPE)) {
          SMF(L.sums, k) = SMF(L.sums, k) + val;
          SMF(L.stepr, k) = per_dt;
        }
      }
      if (sub == 0) SMF(L.rew, 0) = total;
    }
    __syncwarp();
    if (ph & RL_PHASE_COMMAND) {
      const bool skip = (ph & RL_PHASE_SKIP_DONE_ENVS) && done_any;
      command_update<E>(sm, L, S, a, e, env, c, (sub == 0) && !skip);
      __syncwarp();
    }
    if (ph & RL_PHASE_OBS) {
#pragma unroll
      for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
        if (a.out.obs[g] != nullptr && S.obs[g].dim > 0) obs_group<E, LPE>(sm, L, S, a, g, e, sub, env, c);
    }
    __syncthreads();

    // ---- store phase ----------------------------------------------------------------------------
    if (ph & RL_PHASE_OBS) {
      fence_proxy_async();
      __syncthreads();
#pragma unroll
      for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
        if (a.out.obs[g] == nullptr || S.obs[g].dim <= 0) continue;
        const int D = S.obs[g].dim;
        const bool bulk = full && a.out.obs_pitch[g] == D && bulk_ok(a.out.obs[g], D, env0, E);
        if (bulk) {
          if (tid == 0) bulk_s2g(a.out.obs[g] + (size_t)env0 * D, sm + L.obs[g], (uint32_t)(E * D * 4));
        } else {
          for (int i = tid; i < E * D; i += NT) {
            const int el = i / D, col = i % D;
            if (el < nvalid) {
              const long long ev = ids ? (long long)ids[env0 + el] : (long long)(env0 + el);
              a.out.obs[g][ev * a.out.obs_pitch[g] + col] = sm[L.obs[g] + el * L.obs_pitch[g] + col];
            }
          }
        }
      }
      if (tid == 0) bulk_commit();
    }
    if (ph & RL_PHASE_DONES) {
      store_soa<E, int32_t>(sm, L.eplen, a.mdp.episode_length, 1, env0, nvalid, ids, tid, NT);
      if (tid < nvalid) {
        const int fl = __float_as_int(sm[L.flags + tid]);
        const long long ev = ids ? (long long)ids[env0 + tid] : (long long)(env0 + tid);
        if (a.out.done_bits) a.out.done_bits[ev] = (uint8_t)(fl & 0xff);
        if (a.out.terminated) a.out.terminated[ev] = (uint8_t)((fl >> 8) & 1);
        if (a.out.truncated) a.out.truncated[ev] = (uint8_t)((fl >> 9) & 1);
      }
    }
    if (ph & RL_PHASE_REWARDS) {
      RlField fr{a.out.reward, 1, 0};
      store_soa<E, float>(sm, L.rew, fr, 1, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.sums, a.mdp.episode_sums, K, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.stepr, a.out.step_reward, K, env0, nvalid, ids, tid, NT);
    }
    if (do_reset) {
      store_soa<E, float>(sm, L.sums, a.mdp.episode_sums, K, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.act, a.mdp.action, A, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.pact, a.mdp.prev_action, A, env0, nvalid, ids, tid, NT);
      store_soa<E, int32_t>(sm, L.eplen, a.mdp.episode_length, 1, env0, nvalid, ids, tid, NT);
    }
    if ((ph & RL_PHASE_COMMAND) || do_reset) {
      store_soa<E, float>(sm, L.cmd, a.mdp.command, 3, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.head, a.mdp.heading_target, 1, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.tleft, a.mdp.time_left, 1, env0, nvalid, ids, tid, NT);
      store_soa<E, uint8_t>(sm, L.ishead, a.mdp.is_heading_env, 1, env0, nvalid, ids, tid, NT);
      store_soa<E, uint8_t>(sm, L.isstand, a.mdp.is_standing_env, 1, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.mxy, a.mdp.metric_error_vel_xy, 1, env0, nvalid, ids, tid, NT);
      store_soa<E, float>(sm, L.myaw, a.mdp.metric_error_vel_yaw, 1, env0, nvalid, ids, tid, NT);
    }
  }

  // ---- ordered compaction of reset ids (ManagerBasedRLEnv.step: reset_buf.nonzero() [IL]) -----------
  if ((ph & RL_PHASE_COMPACT) && (ph & RL_PHASE_DONES) && !a.has_ids) {
    // per-CTA bit mask of done envs (E <= 32)
    if (nvalid > 0 && tid < E) {
      const int fl = (tid < nvalid) ? __float_as_int(sm[L.flags + tid]) : 0;
      const unsigned m = __ballot_sync(E == 32 ? 0xffffffffu : ((1u << E) - 1u), (fl >> 8) & 3);
      if (tid == 0) a.cta_mask[blockIdx.x] = m;
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(a.ticket, 1u);
      s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      // gridDim.x masks; thread i owns a contiguous run of CTAs -> ids come out ascending
      int* s_cnt = reinterpret_cast<int*>(sm);  // tile data is dead for this CTA only after its stores:
      if (ph & RL_PHASE_OBS) { if (tid == 0) bulk_wait_read0(); }
      __syncthreads();
      const int G = gridDim.x;
      const int per = (G + NT - 1) / NT;
      const int g0 = tid * per, g1 = min(G, g0 + per);
      int cnt = 0;
      for (int g = g0; g < g1; ++g) cnt += __popc(__ldcg(a.cta_mask + g));
      s_cnt[tid] = cnt;
      __syncthreads();
      if (tid == 0) {
        int run = 0;
        for (int i = 0; i < NT; ++i) { const int v = s_cnt[i]; s_cnt[i] = run; run += v; }
        if (a.out.n_reset) *a.out.n_reset = run;
        *a.ticket = 0u;
      }
      __syncthreads();
      int pos = s_cnt[tid];
      if (a.out.reset_ids)
        for (int g = g0; g < g1; ++g) {
          unsigned m = __ldcg(a.cta_mask + g);
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            a.out.reset_ids[pos++] = g * E + b;
          }
        }
      return;
    }
  }
  if (do_reset && nvalid > 0) {
    const int n_cta = (n_total + E - 1) / E;
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(a.ticket, 1u);
      s_last = (prev == (unsigned)(n_cta - 1));
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (tid < K + RL_MAX_DONE_TERMS + 2) {
        float tot = 0.f;
        for (int g = 0; g < n_cta; ++g) tot += __ldcg(a.log_partials + (size_t)g * RL_LOG_STRIDE + tid);
        const RlResetLog& lg = a.out.reset_log;
        if (tid < K) { if (lg.episode_sum_mean) lg.episode_sum_mean[tid] = tot / (float)n_total; }
        else if (tid < K + RL_MAX_DONE_TERMS) { if (lg.done_term_count) lg.done_term_count[tid - K] = tot; }
        else if (lg.metric_mean) lg.metric_mean[tid - K - RL_MAX_DONE_TERMS] = tot / (float)n_total;
      }
      if (tid == 0) *a.ticket = 0u;
    }
  }
  if ((ph & RL_PHASE_OBS) && nvalid > 0 && tid == 0) bulk_wait_read0();  // smem must outlive the bulk reads
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
  int E, LPE;
  Layout L;
  unsigned int* ticket;
  uint32_t* cta_mask;
  float* log_partials;
  int cta_mask_cap;
  RlRewardTerm* adhoc_dev;
  int sm_count;
  int use_pdl;
};

namespace {

int align_up(int v, int a) { return (v + a - 1) / a * a; }

Layout make_layout(const RlStepSpec& s, int E) {
  Layout L;
  memset(&L, 0, sizeof(L));
  int w = 0;
  auto take = [&](int n) { int o = w * E; w += n; return o; };
  const int J = s.num_joints, A = s.action.n_actions, K = s.num_reward_terms;
  L.root_pos = take(3); L.quat = take(4); L.lin_vel = take(3); L.ang_vel = take(3);
  L.jpos = take(J); L.jvel = take(J); L.jacc = take(J); L.jtau = take(J);
  L.act = take(A); L.pact = take(A);
  L.cmd = take(3); L.head = take(1); L.tleft = take(1); L.ishead = take(1); L.isstand = take(1);
  L.mxy = take(1); L.myaw = take(1); L.eplen = take(1);
  L.sums = take(K);
  L.cair = take(s.num_time_bodies); L.lair = take(s.num_time_bodies);
  L.ccon = take(s.num_time_bodies); L.lcon = take(s.num_time_bodies);
  L.bpos = take(s.num_asset_bodies * 3); L.bvel = take(s.num_asset_bodies * 3);
  L.raypos = take(1);
  L.cmdu = take(RL_NUM_CMD_UNIFORMS);
  L.bmax = take(s.num_hist_bodies);
  L.rew = take(1); L.flags = take(1); L.stepr = take(K);
  L.soa_words = w;
  int off = align_up(w * E, 32);  // 128-byte aligned AoS sections (bulk copies need 16 B)
  L.hist_pitch = s.hist_len * s.num_hist_bodies * 3;
  L.hist = off; off = align_up(off + E * L.hist_pitch, 32);
  L.rays_pitch = s.num_rays;
  L.rays = off; off = align_up(off + E * L.rays_pitch, 32);
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
    L.obs_pitch[g] = s.obs[g].dim;
    L.obs[g] = off; off = align_up(off + E * L.obs_pitch[g], 32);
  }
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
    L.obsu[g] = off; off = align_up(off + E * L.obs_pitch[g], 32);
  }
  L.total_words = off;
  return L;
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
    }
    if (dim != G.dim) return fail(RL_EINVAL, "obs group %s%lld: dim %lld != sum of term dims", "", g, G.dim);
  }
  return RL_OK;
}

template <int E, int LPE, int MODE>
int launch_step(RlCtx* ctx, const KArgs& a, int n_items, cudaStream_t st) {
  const size_t smem = (size_t)ctx->L.total_words * 4;
  static thread_local int configured_device = -1;
  static thread_local size_t configured_smem = 0;
  if (configured_device != ctx->device || configured_smem < smem) {
    CUDA_TRY(cudaFuncSetAttribute(mdp_step_kernel<E, LPE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_device = ctx->device; configured_smem = smem;
  }
  const int grid = (n_items + E - 1) / E;
  if (grid <= 0) return RL_OK;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(E * LPE); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = a.use_pdl ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, mdp_step_kernel<E, LPE, MODE>, a));
  return RL_OK;
}

template <int MODE>
int dispatch_step(RlCtx* ctx, const KArgs& a, int n_items, cudaStream_t st) {
  const int key = ctx->E * 100 + ctx->LPE;
  switch (key) {
    case 808: return launch_step<8, 8, MODE>(ctx, a, n_items, st);
    case 816: return launch_step<8, 16, MODE>(ctx, a, n_items, st);
    case 1604: return launch_step<16, 4, MODE>(ctx, a, n_items, st);
    case 1608: return launch_step<16, 8, MODE>(ctx, a, n_items, st);
    case 3204: return launch_step<32, 4, MODE>(ctx, a, n_items, st);
    case 3208: return launch_step<32, 8, MODE>(ctx, a, n_items, st);
    default: return fail(RL_EINVAL, "unsupported launch config E=%s%lld LPE=%lld", "", ctx->E, ctx->LPE);
  }
}

int ensure_scratch(RlCtx* ctx, int grid) {
  if (grid <= ctx->cta_mask_cap) return RL_OK;
  if (ctx->cta_mask) CUDA_TRY(cudaFree(ctx->cta_mask));
  if (ctx->log_partials) CUDA_TRY(cudaFree(ctx->log_partials));
  ctx->cta_mask = nullptr;
  ctx->log_partials = nullptr;
  const int cap = grid * 2 + 64;
  CUDA_TRY(cudaMalloc(&ctx->cta_mask, sizeof(uint32_t) * cap));
  CUDA_TRY(cudaMemset(ctx->cta_mask, 0, sizeof(uint32_t) * cap));
  CUDA_TRY(cudaMalloc(&ctx->log_partials, sizeof(float) * (size_t)cap * RL_LOG_STRIDE));
  CUDA_TRY(cudaMemset(ctx->log_partials, 0, sizeof(float) * (size_t)cap * RL_LOG_STRIDE));
  ctx->cta_mask_cap = cap;
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
  ctx->E = 16;
  ctx->LPE = 8;
  ctx->L = make_layout(ctx->spec, ctx->E);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  if ((size_t)ctx->L.total_words * 4 > (size_t)prop.sharedMemPerBlockOptin) {
    ctx->E = 8;
    ctx->L = make_layout(ctx->spec, ctx->E);
  }
  CUDA_TRY(cudaMemcpyToSymbol(c_spec, spec, sizeof(RlStepSpec), sizeof(RlStepSpec) * slot));
  CUDA_TRY(cudaMalloc(&ctx->ticket, sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(ctx->ticket, 0, sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc(&ctx->adhoc_dev, sizeof(RlRewardTerm) * 64));
  rc = ensure_scratch(ctx, 4096);
  if (rc != RL_OK) return rc;
  g_slots[device][slot] = true;
  *out = ctx;
  return RL_OK;
}

void rl_ctx_destroy(RlCtx* ctx) {
  if (!ctx) return;
  DeviceGuard guard(ctx->device);
  if (ctx->ticket) cudaFree(ctx->ticket);
  if (ctx->cta_mask) cudaFree(ctx->cta_mask);
  if (ctx->log_partials) cudaFree(ctx->log_partials);
  if (ctx->adhoc_dev) cudaFree(ctx->adhoc_dev);
  g_slots[ctx->device][ctx->slot] = false;
  delete ctx;
}

int rl_ctx_set_launch_config(RlCtx* ctx, int envs_per_cta, int lanes_per_env) {
  if (!ctx) return fail(RL_EINVAL, "null ctx%s", "");
  const int E = envs_per_cta > 0 ? envs_per_cta : 16;
  const int LPE = lanes_per_env > 0 ? lanes_per_env : 8;
  const int key = E * 100 + LPE;
  const int okeys[] = {808, 816, 1604, 1608, 3204, 3208};
  bool ok = false;
  for (int k : okeys) ok = ok || (k == key);
  if (!ok) return fail(RL_EINVAL, "unsupported launch config E=%s%lld LPE=%lld", "", E, LPE);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
  Layout L = make_layout(ctx->spec, E);
  if ((size_t)L.total_words * 4 > (size_t)prop.sharedMemPerBlockOptin)
    return fail(RL_EINVAL, "launch config E=%s%lld needs %lld bytes of shared memory", "", E, (long long)L.total_words * 4);
  ctx->E = E; ctx->LPE = LPE; ctx->L = L;
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
  // required pointers per phase
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
  if (phases & RL_PHASE_COMMAND) {
    if (!mdp->heading_target.ptr || !mdp->time_left.ptr || !mdp->is_heading_env.ptr || !mdp->is_standing_env.ptr ||
        !mdp->metric_error_vel_xy.ptr || !mdp->metric_error_vel_yaw.ptr)
      return fail(RL_EINVAL, "rl_step: command state fields required for the command phase%s", "");
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
  if ((phases & RL_PHASE_RESET) && (!mdp->episode_sums.ptr || !mdp->prev_action.ptr || !mdp->heading_target.ptr || !mdp->time_left.ptr ||
      !mdp->is_heading_env.ptr || !mdp->is_standing_env.ptr || !mdp->metric_error_vel_xy.ptr || !mdp->metric_error_vel_yaw.ptr))
    return fail(RL_EINVAL, "rl_step: RL_PHASE_RESET needs every RlMdpState field%s", "");
  if (env_ids && (phases & RL_PHASE_COMPACT)) return fail(RL_EINVAL, "rl_step: compaction is not available on an env_ids subset%s", "");
  if ((phases & RL_PHASE_COMPACT) && !(phases & RL_PHASE_DONES)) return fail(RL_EINVAL, "rl_step: COMPACT needs DONES%s", "");
  DeviceGuard guard(ctx->device);
  const int grid = (int)((num_envs + ctx->E - 1) / ctx->E);
  if (phases & (RL_PHASE_COMPACT | RL_PHASE_RESET)) {
    if (grid > ctx->cta_mask_cap) {
      cudaStreamCaptureStatus cs;
      CUDA_TRY(cudaStreamIsCapturing((cudaStream_t)stream, &cs));
      if (cs != cudaStreamCaptureStatusNone) return fail(RL_EINVAL, "rl_step: first call at this num_envs must happen outside stream capture%s", "");
      int rc = ensure_scratch(ctx, grid);
      if (rc != RL_OK) return rc;
    }
  }
  KArgs a;
  memset(&a, 0, sizeof(a));
  a.N = (int)num_envs; a.slot = ctx->slot; a.phases = phases; a.has_ids = env_ids != nullptr;
  a.st = *state; a.mdp = *mdp; a.out = *out; a.rnd = *rnd;
  a.env_ids = env_ids; a.n_env_ids = n_env_ids; a.L = ctx->L;
  a.ticket = ctx->ticket; a.cta_mask = ctx->cta_mask; a.log_partials = ctx->log_partials; a.use_pdl = ctx->use_pdl;
  return dispatch_step<0>(ctx, a, (int)num_envs, (cudaStream_t)stream);
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
  const int grid = (int)((num_envs + ctx->E - 1) / ctx->E);
  if (grid > ctx->cta_mask_cap) {
    cudaStreamCaptureStatus cs;
    CUDA_TRY(cudaStreamIsCapturing((cudaStream_t)stream, &cs));
    if (cs != cudaStreamCaptureStatusNone) return fail(RL_EINVAL, "rl_reset_envs: first call at this num_envs must happen outside stream capture%s", "");
    int rc = ensure_scratch(ctx, grid);
    if (rc != RL_OK) return rc;
  }
  KArgs a;
  memset(&a, 0, sizeof(a));
  a.N = (int)num_envs; a.slot = ctx->slot; a.phases = RL_PHASE_RESET; a.has_ids = 1;
  a.mdp = *mdp; a.rnd = *rnd; a.out.done_bits = const_cast<uint8_t*>(done_bits);
  if (log) a.out.reset_log = *log;
  a.env_ids = env_ids; a.n_env_ids = n_env_ids; a.L = ctx->L;
  a.ticket = ctx->ticket; a.cta_mask = ctx->cta_mask; a.log_partials = ctx->log_partials; a.use_pdl = ctx->use_pdl;
  return dispatch_step<0>(ctx, a, (int)num_envs, (cudaStream_t)stream);
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
  memset(&a, 0, sizeof(a));
  a.N = (int)num_envs; a.slot = ctx->slot; a.phases = 0; a.has_ids = 0;
  a.st = *state; a.mdp = *mdp; a.L = ctx->L;
  a.adhoc = dev; a.ext_terminated = terminated; a.term_out = out; a.use_pdl = 0;
  return dispatch_step<1>(ctx, a, (int)num_envs, st);
}

}  // extern "C"
