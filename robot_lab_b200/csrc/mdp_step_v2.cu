// mdp_step_v2.cu - the two launches of an env step as thread-block-cluster kernels for sm_100a (B200).
//
// What profiles/r1_summary.md measured on the general kernel (csrc/mdp_step.cu) at 4096 envs: every SM runs one tile
// and executes 35-41 KB of distinct code exactly once per launch; instruction supply, a 590-instruction load prologue
// and the serial phases of a tile - not bytes - are the time. This file is the answer to that for the two launch
// kinds an env step consists of (ManagerBasedRLEnv.step [IL]):
//
//   RL_V2_PRE   = DONES | REWARDS | COMPACT          (termination terms, reward terms, ordered reset ids)
//   RL_V2_POST  = RESET (masked) | COMMAND | OBS     (manager reset of the done envs, command.compute, both obs groups)
//
//   * a configuration is (C, G, NW): a cluster of C CTAs owns G tiles (32 G consecutive envs); CTA r of the cluster is
//     ROLE r and evaluates the r-th share of the task list for all G tiles with NW warps = G tiles x W task slots; warp
//     (tile, slot) runs the tasks of bin role*W + slot for its tile, lane = env. What production launches use is
//     C = G = 1 - one tile per CTA, 16 warps up to 16384 envs, 8 beyond; the cluster forms (2 x 2, 4 x 4: an SM executes
//     1/C of the term code, every instruction serves G warps) are compiled, tested and were measured: not faster
//     (profiles/r2_summary.md);
//   * load: one 2-D TMA tensor-map copy per SoA field ({32 G envs} x {components} box of the [C, N] tensor, SASS
//     UTMALDG), issued by lane 0 of EVERY warp (one thread issuing all of them serialised the descriptors' first-touch
//     misses); the contact-force rows are not staged at all;
//   * contact-force norms (max over the history of |F_b|) are computed once per env by a prepass that streams the force
//     rows from global memory while the copies are in flight, and cached in the record (every consumer used to recompute
//     them in its own serial chain);
//   * results leave from registers: lane = env makes every SoA result row a coalesced 128-byte store by the warp that
//     produced it - there is no store phase; the weighted term values meet in role 0's shared memory (in a cluster:
//     through DSMEM st.shared::cluster and one cluster barrier), where one warp per tile adds them up in manager order -
//     the same summation order as the general kernel and the reference (RewardManager.compute [IL]);
//   * the ordered reset-id list has no launch-wide tail: decoupled look-back over one status word per tile
//     (lookback_publish / lookback_resolve); the logging means of the reset keep one (a release ticket per LOG task, an
//     acquiring load by the thread that drew the last one);
//   * the height-scan observation (187 of the 235 critic columns) never enters shared memory: every warp streams env
//     rows global -> registers -> global with lanes = columns while the TMA loads are still in flight.
//
// Same term functions (csrc/mdp_terms.cuh), same operand order as the general kernel: all results are bit-identical to
// it (tests/test_gpu_v2_parity.py). Launches that do not qualify (ragged env counts, env-id lists, strided IsaacLab
// tensors, a spec that is not baked) run the general kernel.

#ifndef RL_V2_UNROLL
#define RL_V2_UNROLL 0   // measured (profiles/r2_summary.md): unrolling the term loops costs more cold code than it saves
#endif
#if RL_V2_UNROLL
#define RL_TERM_LOOP _Pragma("unroll")   // see csrc/mdp_terms.cuh: the term loops unroll against the baked spec
#endif
#ifndef RL_V2_FEW_UNROLL
#define RL_V2_FEW_UNROLL 1   // the short loops (feet, contact-mask bodies) unroll: they were the critical path of a tile
#endif
#if RL_V2_FEW_UNROLL
#define RL_FEW_LOOP _Pragma("unroll")
#endif
#include "mdp_ctx.h"

#include <algorithm>

#include <cuda.h>

#ifndef RL_V2_DEV_ONE
#define RL_V2_DEV_ONE 0   // development builds: compile the cluster kernels for Go2-rough only (seconds instead of minutes)
#endif

#ifndef RL_V2_STAMPS
#define RL_V2_STAMPS 0   // build variant "stamps": clock64 / globaltimer stamps per CTA and warp into KArgs::dbg (tools/v2_timeline.py)
#endif
#if RL_V2_STAMPS
#define V2_STAMP(slot) do { if (a.k.dbg > reinterpret_cast<long long*>(2) && lane == 0) a.k.dbg[(size_t)blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#define V2_STAMP_T0(slot) do { if (a.k.dbg > reinterpret_cast<long long*>(2) && tid == 0) a.k.dbg[(size_t)blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#define V2_GTIME(slot) do { if (a.k.dbg > reinterpret_cast<long long*>(2) && tid == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); a.k.dbg[(size_t)blockIdx.x * 64 + (slot)] = (long long)g_; } } while (0)
#define V2_DBG_ROW (a.k.dbg > reinterpret_cast<long long*>(2) ? a.k.dbg + (size_t)blockIdx.x * 64 : nullptr)
#define V2_GTIME_L(k_, slot) do { if ((k_).dbg > reinterpret_cast<long long*>(2) && (threadIdx.x & 31) == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); (k_).dbg[(size_t)blockIdx.x * 64 + (slot)] = (long long)g_; } } while (0)
#define V2_GTIME_K(k_, slot) do { if ((k_).dbg > reinterpret_cast<long long*>(2) && threadIdx.x == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); (k_).dbg[(size_t)blockIdx.x * 64 + (slot)] = (long long)g_; } } while (0)
#else
#define V2_STAMP(slot) do { } while (0)
#define V2_STAMP_T0(slot) do { } while (0)
#define V2_GTIME(slot) do { } while (0)
#define V2_DBG_ROW nullptr
#define V2_GTIME_L(k_, slot) do { } while (0)
#define V2_GTIME_K(k_, slot) do { } while (0)
#endif

namespace {

constexpr int kLogWarps = 16;   // the logging reduction keeps the general kernel's order: 16 strided partial sums, then their sum

// ---------------------------------------------------------------------------------------------------
// Which input fields a launch kind stages (the sets fill_args() of the general kernel uses for the same phases)
// ---------------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t v2_field_mask(const RlStepSpec& s, int kind) {
  uint32_t m = (1u << IF_ROOT_POS) | (1u << IF_QUAT) | (1u << IF_LIN_VEL) | (1u << IF_ANG_VEL) | (1u << IF_JPOS) |
               (1u << IF_JVEL) | (1u << IF_CMD) | (1u << IF_EPLEN) | (1u << IF_ACT);
  if (kind == RL_V2_PRE) {
    // (the episode sums are not staged: the task that updates sum k reads its row from global memory when it starts and
    //  writes it when it ends - the same bytes, 21 fewer record rows: a fifth resident CTA per SM at 8 warps per tile)
    m |= (1u << IF_PACT) | (1u << IF_JACC) | (1u << IF_JTAU) | (1u << IF_CAIR) | (1u << IF_LAIR) | (1u << IF_CCON) |
         (1u << IF_LCON) | (1u << IF_BPOS) | (1u << IF_BVEL);
  } else {
    m |= (1u << IF_HEAD) | (1u << IF_TLEFT) | (1u << IF_MXY) | (1u << IF_MYAW) | (1u << IF_CMDU) | (1u << IF_SUMS);
  }
  for (int f = 0; f < IF_COUNT; ++f)
    if (in_field_ncomp(s, f) <= 0) m &= ~(1u << f);
  return m;
}

// index of the group's height-scan term (must be its last term: checked by v2_spec_ok) or -1
__host__ __device__ constexpr int scan_term_of(const RlObsGroup& G) {
  for (int t = 0; t < G.n_terms; ++t)
    if (G.terms[t].type == RL_OBS_HEIGHT_SCAN) return t;
  return -1;
}
// columns of a group row that are assembled per env in shared memory (everything but the streamed height scan)
__host__ __device__ constexpr int env_cols_of(const RlObsGroup& G) {
  const int st = scan_term_of(G);
  return st < 0 ? G.dim : G.dim - G.terms[st].dim;
}

// what the cluster kernels cover (everything else runs the general kernel)
__host__ __device__ constexpr bool v2_spec_ok(const RlStepSpec& s) {
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
    const RlObsGroup& G = s.obs[g];
    int scans = 0;
    for (int t = 0; t < G.n_terms; ++t)
      if (G.terms[t].type == RL_OBS_HEIGHT_SCAN) {
        ++scans;
        if (t != G.n_terms - 1) return false;                          // streamed columns must end the row
        if (G.terms[t].has_noise && G.enable_corruption) return false;   // noise is generated lane = env
      }
    if (scans > 1) return false;
  }
  if (s.num_hist_bodies > 0 && s.hist_len <= 0) return false;
  for (int k = 0; k < s.num_reward_terms; ++k)   // feet_stumble reads raw force rows (the newest sample), which are not staged here
    if (s.rewards[k].type == RL_REW_FEET_STUMBLE && s.rewards[k].weight != 0.f) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------------
// Record layout of a CTA: 32 G envs; the term functions address it through the same Layout members as the general
// kernel (SoA word w of env e at sm[w*E + e]).
// ---------------------------------------------------------------------------------------------------
__host__ __device__ constexpr Layout make_layout_v2(const RlStepSpec& s, const int E, const int kind, const int NW) {
  Layout L{};
  L.E = E;
  int w = 0;
  auto take = [&w](int n) { int o = w; w += n; return o; };
  const int J = s.num_joints, A = s.action.n_actions, K = s.num_reward_terms;
  L.A = A; L.J = J; L.K = K;
  const uint32_t mask = v2_field_mask(s, kind);
  int iw[IF_COUNT] = {};
  for (int f = 0; f < IF_COUNT; ++f) iw[f] = ((mask >> f) & 1u) ? take(in_field_ncomp(s, f)) : 0;
  L.root_pos = iw[IF_ROOT_POS] * E; L.quat = iw[IF_QUAT] * E; L.lin_vel = iw[IF_LIN_VEL] * E; L.ang_vel = iw[IF_ANG_VEL] * E;
  L.jpos = iw[IF_JPOS] * E; L.jvel = iw[IF_JVEL] * E; L.jacc = iw[IF_JACC] * E; L.jtau = iw[IF_JTAU] * E;
  L.cair = iw[IF_CAIR] * E; L.lair = iw[IF_LAIR] * E; L.ccon = iw[IF_CCON] * E; L.lcon = iw[IF_LCON] * E;
  L.bpos = iw[IF_BPOS] * E; L.bvel = iw[IF_BVEL] * E;
  L.cmd = iw[IF_CMD] * E; L.head = iw[IF_HEAD] * E; L.tleft = iw[IF_TLEFT] * E;
  L.mxy = iw[IF_MXY] * E; L.myaw = iw[IF_MYAW] * E; L.eplen = iw[IF_EPLEN] * E;
  L.sums = iw[IF_SUMS] * E; L.cmdu = iw[IF_CMDU] * E; L.act = iw[IF_ACT] * E; L.pact = iw[IF_PACT] * E;
  L.w_sums = iw[IF_SUMS];
  if (kind == RL_V2_POST) {
    L.ishead = take(1) * E; L.isstand = take(1) * E; L.rmask = take(1) * E;
    L.cmdn = take(3) * E; L.epnew = take(1) * E;
  }
  L.flags = take(1) * E;
  if (kind == RL_V2_PRE) {
    L.termv = take(K) * E;                      // role 0: weighted value of every term (DSMEM target)
    L.hnorm = take(s.num_hist_bodies) * E;      // cached contact-force norms
  }
  L.soa_words = w;
  int off = align_up(w * E, 32);
  L.cj = off; off = align_up(off + 5 * J, 32);
  if (kind == RL_V2_PRE) {
    // the contact-force rows are NOT staged: the norm prepass streams them from global memory; per warp a scratch row of
    // envs-per-warp * hist_len * bodies squared norms (hist = scratch base, hist_pitch = words per warp)
    L.hist_pitch = align_up((E / NW) * s.hist_len * s.num_hist_bodies, 32);   // all of a warp's envs at once
    L.hist = off; off = align_up(off + NW * L.hist_pitch, 32);
  } else {
    L.obs_pitch0 = odd_pitch(env_cols_of(s.obs[0])); L.obs_pitch1 = odd_pitch(env_cols_of(s.obs[1]));
    L.obs0 = off; off = align_up(off + E * L.obs_pitch0, 32);
    L.obs1 = off; off = align_up(off + E * L.obs_pitch1, 32);
  }
  L.total_words = off;
  return L;
}
// record word a field is staged at (host side of the TMA issue loop)
__host__ __device__ constexpr int v2_field_word(const RlStepSpec& s, int kind, int f) {
  const uint32_t mask = v2_field_mask(s, kind);
  int w = 0;
  for (int i = 0; i < f; ++i)
    if ((mask >> i) & 1u) w += in_field_ncomp(s, i);
  return w;
}

// ---------------------------------------------------------------------------------------------------
// Schedule: tasks -> bins (bin = role * W + slot). Longest-processing-time greedy with two placement rules:
//   PRE : the termination terms and every consumer of contact-force norms live in role 0 (it stages the force rows,
//         runs the norm prepass and owns the final sum); everything else is balanced over all bins;
//   POST: the command task lives in role 0, slot 0; the per-env columns of observation group g are assembled by role
//         owner(g) (policy: role 1 if it exists, critic: role 2 if it exists, else role 0).
// ---------------------------------------------------------------------------------------------------
struct Sched2 {
  int n;
  Task t[RL_MAX_TASKS];
  uint64_t late;          // reward terms finished by the final sum (is_terminated)
  uint32_t hist_roles;    // roles that stage the contact-force rows and run the norm prepass
  int obs_owner[RL_NUM_OBS_GROUPS];
  int log_parts;          // POST: number of TK_LOG tasks per tile (= ticket arrivals per tile)
};

__host__ __device__ constexpr bool term_uses_hist(const RlRewardTerm& t) {
  return t.type == RL_REW_UNDESIRED_CONTACTS || t.type == RL_REW_CONTACT_FORCES || t.type == RL_REW_FEET_SLIDE ||
         t.type == RL_REW_FEET_STUMBLE;
}
__host__ __device__ constexpr bool dones_use_hist(const RlStepSpec& s) {
  for (int d = 0; d < s.num_done_terms; ++d)
    if (s.dones[d].type == RL_DONE_ILLEGAL_CONTACT) return true;
  return false;
}

__host__ __device__ constexpr Sched2 make_schedule_v2(const RlStepSpec& s, int kind, int C, int W, int G, int NW) {
  Sched2 sc{};
  int cost[RL_MAX_TASKS] = {};
  int lo_bin[RL_MAX_TASKS] = {}, hi_bin[RL_MAX_TASKS] = {};   // bins a task may go to: [lo, hi)
  const int BINS = C * W;
  int n = 0;
  int load[64] = {};
  sc.obs_owner[0] = C > 1 ? 1 : 0;
  sc.obs_owner[1] = C > 2 ? 2 : 0;
  if (kind == RL_V2_PRE) {
    bool any_hist = dones_use_hist(s);
    // the termination task: not in bin 0 (the final sum's warp) when there is a choice - its look-back overlaps the final sum
    sc.t[n] = Task{TK_DONES, 0, 0, 0, 0, 0, 0, 0}; cost[n] = 120; lo_bin[n] = W > 1 ? 1 : 0; hi_bin[n] = W; ++n;
    for (int k = 0; k < s.num_reward_terms; ++k) {
      const RlRewardTerm& t = s.rewards[k];
      if (t.weight == 0.f) continue;
      if (t.type == RL_REW_IS_TERMINATED) { sc.late |= 1ull << k; continue; }
      const bool body_sum = (t.type == RL_REW_UNDESIRED_CONTACTS || t.type == RL_REW_CONTACT_FORCES);
      sc.t[n] = Task{TK_REWARD, (uint8_t)k, 0, 0, 0, 64, 1, 0};
      cost[n] = 60 + reward_cost(t, s, body_sum ? popc64(t.body_mask) : t.n_idx, true);
      const bool h = term_uses_hist(t) && s.num_hist_bodies > 0;
      any_hist = any_hist || h;
      lo_bin[n] = 0; hi_bin[n] = h ? W : BINS;
      ++n;
    }
    if (any_hist) {
      sc.hist_roles = 1u;
      // what the prepass costs every warp of role 0 before its tasks start
      const int pre = ((G * s.num_hist_bodies + NW - 1) / NW) * (s.hist_len * 18 + 12) + 40;
      for (int b = 0; b < W; ++b) load[b] = pre;
    }
    for (int b = 0; b < W; ++b) load[b] += 80;   // role 0 also runs the final sum and the compaction tail
  } else {
    {
      int c = 520;  // manager reset bookkeeping + command update + heading control
      for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
        for (int ti = 0; ti < s.obs[g].n_terms; ++ti)
          if (s.obs[g].terms[ti].type == RL_OBS_GENERATED_COMMANDS) c += 40;
      sc.t[n] = Task{TK_COMMAND, 0, 0, 0, 0, 0, 0, 0}; cost[n] = c; lo_bin[n] = 0; hi_bin[n] = 1; ++n;
      // logging reductions of the reset + zeroing of the reset envs' sums / stored actions, in up to four parts (part p
      // takes the quantities q = p mod parts): any warp but the command's
      const int parts = BINS > 4 ? 4 : (BINS > 1 ? BINS - 1 : 1);
      sc.log_parts = parts;
      for (int p = 0; p < parts; ++p) {
        sc.t[n] = Task{TK_LOG, 0, (uint8_t)p, 0, 0, 0, (uint16_t)parts, 0};
        cost[n] = 120 + 40 * ((s.num_reward_terms + RL_MAX_DONE_TERMS + 2 + parts - 1) / parts);
        lo_bin[n] = BINS > 1 ? 1 : 0; hi_bin[n] = BINS; ++n;
      }
    }
    for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
      int col0 = 0;
      for (int ti = 0; ti < s.obs[g].n_terms; ++ti) {
        const RlObsTerm& o = s.obs[g].terms[ti];
        const int per_col = 8 + ((o.has_noise && s.obs[g].enable_corruption) ? 24 : 0);
        if (o.type != RL_OBS_GENERATED_COMMANDS && o.type != RL_OBS_HEIGHT_SCAN) {
          constexpr int chunk = 32;   // multiples of 4 (one Philox block = 4 columns)
          for (int lo = 0; lo < o.dim; lo += chunk) {
            const int hi = (lo + chunk < o.dim) ? lo + chunk : o.dim;
            if (n >= RL_MAX_TASKS) break;
            sc.t[n] = Task{TK_OBS, (uint8_t)g, (uint8_t)ti, 0, (uint16_t)lo, (uint16_t)hi, (uint16_t)col0, 0};
            cost[n] = 20 + per_col * (hi - lo);
            lo_bin[n] = sc.obs_owner[g] * W; hi_bin[n] = lo_bin[n] + W;
            ++n;
          }
        }
        col0 += o.dim;
      }
    }
  }
  sc.n = n;
  // longest-processing-time greedy; the constrained tasks first so that the free ones balance around them
  bool done[RL_MAX_TASKS] = {};
  for (int pass = 0; pass < 2; ++pass) {
    for (int it = 0; it < n; ++it) {
      int best = -1;
      for (int i = 0; i < n; ++i) {
        const bool constrained = (hi_bin[i] - lo_bin[i]) < BINS;
        if (done[i] || constrained != (pass == 0)) continue;
        if (best < 0 || cost[i] > cost[best]) best = i;
      }
      if (best < 0) break;
      int w = lo_bin[best];
      for (int j = lo_bin[best]; j < hi_bin[best]; ++j)
        if (load[j] < load[w]) w = j;
      sc.t[best].owner = (uint8_t)w; load[w] += cost[best]; done[best] = true;
    }
  }
  return sc;
}

// ---------------------------------------------------------------------------------------------------
// What a role stages: the union of the input fields its tasks read (a CTA of a 4-role cluster would otherwise pull the
// whole record of 128 envs through its SM for a quarter of the terms).
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kCoreFields = (1u << IF_ROOT_POS) | (1u << IF_QUAT) | (1u << IF_LIN_VEL) | (1u << IF_ANG_VEL) | (1u << IF_CMD);   // make_ctx

__host__ __device__ constexpr uint32_t reward_fields(int type) {
  uint32_t m = kCoreFields;
  switch (type) {
    case RL_REW_JOINT_TORQUES_L2: m |= 1u << IF_JTAU; break;
    case RL_REW_JOINT_VEL_L2: case RL_REW_JOINT_VEL_LIMITS: m |= 1u << IF_JVEL; break;
    case RL_REW_JOINT_ACC_L2: m |= 1u << IF_JACC; break;
    case RL_REW_JOINT_DEVIATION_L1: case RL_REW_JOINT_POS_LIMITS: case RL_REW_STAND_STILL: case RL_REW_JOINT_POS_PENALTY:
    case RL_REW_JOINT_MIRROR: m |= 1u << IF_JPOS; break;
    case RL_REW_JOINT_POWER: m |= (1u << IF_JVEL) | (1u << IF_JTAU); break;
    case RL_REW_ACTION_MIRROR: case RL_REW_ACTION_SYNC: m |= 1u << IF_ACT; break;
    case RL_REW_ACTION_RATE_L2: m |= (1u << IF_ACT) | (1u << IF_PACT); break;
    case RL_REW_FEET_AIR_TIME: m |= (1u << IF_LAIR) | (1u << IF_CCON); break;
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: case RL_REW_FEET_GAIT: m |= (1u << IF_CAIR) | (1u << IF_CCON); break;
    case RL_REW_FEET_AIR_TIME_VARIANCE: m |= (1u << IF_LAIR) | (1u << IF_LCON); break;
    case RL_REW_FEET_CONTACT: case RL_REW_FEET_CONTACT_WITHOUT_CMD: m |= 1u << IF_CCON; break;
    case RL_REW_FEET_SLIDE: m |= 1u << IF_BVEL; break;
    case RL_REW_FEET_HEIGHT: case RL_REW_FEET_HEIGHT_BODY: m |= (1u << IF_BPOS) | (1u << IF_BVEL); break;
    case RL_REW_FEET_DISTANCE_Y_EXP: case RL_REW_FEET_DISTANCE_XY_EXP: m |= 1u << IF_BPOS; break;
    case RL_REW_WHEEL_VEL_PENALTY: m |= (1u << IF_JVEL) | (1u << IF_CAIR); break;
    default: break;
  }
  return m;
}
__host__ __device__ constexpr uint32_t obs_fields(int type) {
  switch (type) {
    case RL_OBS_JOINT_POS_REL: case RL_OBS_JOINT_POS_REL_WITHOUT_WHEEL: return kCoreFields | (1u << IF_JPOS);
    case RL_OBS_JOINT_VEL_REL: return kCoreFields | (1u << IF_JVEL);
    case RL_OBS_LAST_ACTION: return kCoreFields | (1u << IF_ACT);
    case RL_OBS_PHASE: return kCoreFields | (1u << IF_EPLEN);
    default: return kCoreFields;
  }
}
__host__ __device__ constexpr uint32_t role_field_mask(const RlStepSpec& s, const Sched2& sc, int kind, int role, int W) {
  uint32_t m = kCoreFields | (1u << IF_EPLEN);
  if (role == 0) {
    if (kind == RL_V2_POST) m |= (1u << IF_HEAD) | (1u << IF_TLEFT) | (1u << IF_MXY) | (1u << IF_MYAW) | (1u << IF_CMDU);
  }
  for (int i = 0; i < sc.n; ++i) {
    const Task& t = sc.t[i];
    if (t.owner / W != role) continue;
    if (t.kind == TK_REWARD) m |= reward_fields(s.rewards[t.a].type);
    else if (t.kind == TK_OBS) m |= obs_fields(s.obs[t.a].terms[t.b].type);
    else if (t.kind == TK_LOG) m |= (1u << IF_SUMS) | (1u << IF_MXY) | (1u << IF_MYAW);
  }
  return m & v2_field_mask(s, kind);
}

// ---------------------------------------------------------------------------------------------------
// PTX helpers of the cluster kernels
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// programmatic dependent launch: let the next kernel of the stream start its CTAs (their prologue overlaps this kernel),
// and - in the dependent - wait until the predecessor has completed and its memory is visible
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// L2 prefetch of a contiguous span (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// generic address of `p` (a pointer into this CTA's shared memory) in CTA `rank` of the cluster
template <class T> __device__ __forceinline__ T* map_to_rank(T* p, uint32_t rank) {
  uint64_t r;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(r) : "l"(reinterpret_cast<uint64_t>(p)), "r"(rank));
  return reinterpret_cast<T*>(r);
}
// 2-D tiled tensor-map copy global -> shared (SASS UTMALDG), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

struct alignas(64) V2Args {
  KArgs k;                  // the parameter block of the general kernel (pointers, strides, phases, random streams)
  uint32_t role_bytes[8];   // bytes the tensor-map copies of one CTA of role r deliver
  uint8_t role_n[8];                  // fields role r stages (what its tasks read) ...
  uint8_t role_field[8][IF_COUNT];    // ... as a list (a bit mask walked with __fns cost ~100 instructions per copy, 7-8 % of a launch)
  int32_t field_word[IF_COUNT];
  const char* prefetch_rays;   // PRE: the ray-hit rows the post-reset launch will stream (L2 prefetch), or NULL
  uint32_t prefetch_row_bytes;
  // the pre-reset launch's ordered reset ids without a launch-wide tail (scratch owned by the context, see RlV2State):
  unsigned long long* scan_state;   // PRE: one status word per tile (epoch | flag | reset count or prefix): decoupled look-back
  unsigned int* scan_ctl;           // [0] epoch of the look-back
  alignas(64) CUtensorMap tm[IF_COUNT];
};

template <class B, int KIND, int C, int G, int NW>
struct Cfg2 {
  static_assert(NW % G == 0 && NW <= 16, "tiles per CTA must divide the warp count");
  static constexpr int W = NW / G;        // task slots per (role, tile)
  static constexpr int NT = NW * 32;      // threads per CTA
  static constexpr int E = 32 * G;        // envs per cluster
  static constexpr int BINS = C * W;
  static constexpr Layout L = make_layout_v2(B::spec, E, KIND, NW);
  static constexpr Sched2 sched = make_schedule_v2(B::spec, KIND, C, W, G, NW);
  static constexpr Scalars S = scalars_of(B::spec);
  // scalar copies: device code must not odr-use the schedule object itself
  static constexpr int n_tasks = sched.n;
  static constexpr uint64_t late = sched.late;
  static constexpr uint32_t hist_roles = sched.hist_roles;
  static constexpr int obs_owner0 = sched.obs_owner[0], obs_owner1 = sched.obs_owner[1];
  static constexpr int log_parts = sched.log_parts;
};

// all tasks of one bin, in schedule order, behind ONE per-env context (members no task of the bin reads fold away);
// `fin` closes the bin's straight-line code (every warp runs exactly one bin, also an empty one)
template <class B, class CF, int BIN, class F, class FIN>
__device__ __forceinline__ void bin_tasks(const float* sm, const int e, F&& f, FIN&& fin, long long* dbg_row) {
  constexpr Layout L = CF::L;
  const EnvCtx c = make_ctx(sm, L, e);
  static_for(std::make_integer_sequence<int, CF::n_tasks>{}, [&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr Task tk = CF::sched.t[i];
    if constexpr (tk.owner == BIN) {
      static constexpr RlRewardTerm rt = B::spec.rewards[tk.kind == TK_REWARD ? tk.a : 0];
      static constexpr RlObsTerm ot = B::spec.obs[tk.kind == TK_OBS ? tk.a : 0].terms[tk.kind == TK_OBS ? tk.b : 0];
      constexpr bool corrupt = B::spec.obs[tk.kind == TK_OBS ? tk.a : 0].enable_corruption != 0;
#if RL_V2_STAMPS
      const long long t0_ = clock64();
#endif
      f(ic, tk, rt, ot, corrupt, c);
#if RL_V2_STAMPS
      if (dbg_row != nullptr && (threadIdx.x & 31) == 0 && i < 32) dbg_row[32 + i] = clock64() - t0_;   // cycles of task i
#endif
    }
  });
  (void)dbg_row;
  fin();
}
template <class B, class CF, int LO, int HI, class F, class FIN>
__device__ __forceinline__ void dispatch_bin(const int bin, const float* sm, const int e, F&& f, FIN&& fin, long long* dbg_row = nullptr) {
  if constexpr (HI - LO == 1) {
    bin_tasks<B, CF, LO>(sm, e, f, fin, dbg_row);
  } else {
    constexpr int MID = (LO + HI) / 2;
    if (bin < MID) dispatch_bin<B, CF, LO, MID>(bin, sm, e, f, fin, dbg_row); else dispatch_bin<B, CF, MID, HI>(bin, sm, e, f, fin, dbg_row);
  }
}

// The load every CTA starts with: one instruction per field, issued by lane 0 of EVERY warp (issue_loads): the tensor maps live in the kernel's parameter bank, the first touch of each is a ~300-cycle miss, and one thread
// issuing all ~20 of them in a row serialised those misses into 2.9 us of a 4096-env launch (measured: profiles/
// r2_summary.md). Spread over the warps they overlap.
// Thread 0 initialises the mbarrier (one arrival); after the CTA barrier that publishes it, lane 0 of warp 0 makes that
// arrival with the byte count of the record (arrive.expect_tx) and lane 0 of every warp issues its share of the copies.
// A copy of another warp may complete before the expect_tx is performed - the transaction count may run negative; the
// phase cannot complete before the one pending arrival is made. (init and expect_tx on either side of the CTA barrier
// also keeps compute-sanitizer racecheck quiet about the two accesses of thread 0.)
__device__ __forceinline__ void init_load_barrier(uint64_t* bar) {
  mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
template <class CF>
__device__ __forceinline__ void issue_loads(float* sm, const V2Args& a, const uint32_t role, const int env0, uint64_t* bar, const int warp) {
  if (warp == 0) mbar_expect_tx(bar, a.role_bytes[role]);
  const int n = a.role_n[role];
#pragma unroll 1
  for (int i = warp; i < n; i += CF::NT / 32) {
    const int f = a.role_field[role][i];
    tma_load_2d(sm + a.field_word[f] * CF::E, &a.tm[f], env0, 0, bar);
  }
}

// Contact-force norm prepass: max over the history of |F_b| for every body of every env of the CTA, ONCE, straight from
// global memory. A warp takes an env at a time with lanes = (history sample, body) pairs: its 3 * T * B floats are one
// contiguous, coalesced row; the squared norms go through a per-warp scratch row, lanes = (env, body) items take the max
// over the history (same order as hist_max_norm) and ONE IEEE sqrt. All loads of the warp's envs are in flight before the first
// use, and the whole prepass runs while the tensor-map copies of the record are still in flight. Nothing of the force
// history is staged in shared memory (round-2 first cut: a 22 KB bulk copy per tile, half of the record).
template <class CF>
__device__ __forceinline__ void norm_prepass(float* sm, const float* __restrict__ hist_tile, const int warp, const int lane) {
  constexpr Layout L = CF::L;
  constexpr Scalars S = CF::S;
  constexpr int T = S.hist_len, Bh = S.num_hist_bodies, TB = T * Bh, HW = TB * 3;
  constexpr int NWARPS = CF::NT / 32, EPW = CF::E / NWARPS, PASSES = (TB + 31) / 32;
  static_assert(CF::E % NWARPS == 0, "envs of a CTA must divide over its warps");
  float ss[EPW][PASSES];
#pragma unroll
  for (int i = 0; i < EPW; ++i) {
    const float* row = hist_tile + (size_t)(warp + i * NWARPS) * HW;
#pragma unroll
    for (int q = 0; q < PASSES; ++q) {
      const int p = q * 32 + lane;
      ss[i][q] = 0.f;
      if (p < TB) {
        const float f0 = __ldg(row + 3 * p), f1 = __ldg(row + 3 * p + 1), f2 = __ldg(row + 3 * p + 2);
        ss[i][q] = (f0 * f0 + f1 * f1) + f2 * f2;
      }
    }
  }
  float* scr = sm + L.hist + warp * L.hist_pitch;
#pragma unroll
  for (int i = 0; i < EPW; ++i)
#pragma unroll
    for (int q = 0; q < PASSES; ++q)
      if (q * 32 + lane < TB) scr[i * TB + q * 32 + lane] = ss[i][q];
  __syncwarp();
  // lanes = (env of the warp, body) items: every lane busy (bodies alone would leave 19 of 32 lanes idle on a quadruped)
#pragma unroll 1
  for (int it = lane; it < EPW * Bh; it += 32) {
    const int i = it / Bh, b = it - i * Bh;
    const float* sc = scr + i * TB + b;
    float m2 = sc[0];
#pragma unroll
    for (int t = 1; t < T; ++t) m2 = fmaxf(m2, sc[t * Bh]);
    const bool z = (m2 == 0.f);                   // a body without contact: keep the warp off sqrt's special-operand path
    const float r = sqrtf(z ? 1.f : m2);
    sm[L.hnorm + b * CF::E + (warp + i * NWARPS)] = z ? 0.f : r;
  }
}

// reset_buf.nonzero() [IL]: ascending reset ids WITHOUT a launch-wide tail. Every tile publishes the number of its reset
// envs as soon as its termination terms are evaluated (one warp, early in the tile's life) and gets its offset into the
// id list by a decoupled look-back over the tiles in front of it (Merrill & Garland's single-pass scan): a status word is
// (epoch << 32) | (flag << 24) | value with flag 1 = the tile's own count, 2 = inclusive prefix. The whole payload is in
// the one 64-bit word, so relaxed loads and stores are enough. The epoch (a word in device memory, bumped by the LAST
// tile once its own look-back is complete - by then every tile has published, i.e. has read the epoch) makes the words
// of earlier launches invalid without clearing them, also when a captured graph replays the same arguments.
// Forward progress: a tile waits only for tiles with a smaller block index, which the hardware dispatches first (the
// assumption of every single-pass scan); the spin is bounded and traps instead of hanging.
// (First cut of this round: the CTA that drew the last ticket compacted the ids of the whole launch with every other SM
// idle - 1.6 us of a 9.5 us launch at 4096 envs, 5.4 us at 65536; profiles/r2_summary.md.)
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Two halves, both run by the warp of the tile's termination task: publish (as soon as the terms are evaluated) and
// resolve (after the tile's task barrier, next to the final sum: by then the tiles in front have published long ago, at
// 4096 envs - every tile in lock step - as well as in the later waves of a large launch, so the resolve is one round
// trip that overlaps the final sum instead of a spin inside the longest task of the tile).
__device__ __forceinline__ void lookback_publish(unsigned long long* st, const unsigned epoch, const int gt, const unsigned cnt, const int lane) {
  const unsigned long long tag = (unsigned long long)epoch << 32;
  if (lane == 0) st_relaxed_u64(st + gt, tag | ((gt == 0 ? 2ull : 1ull) << 24) | cnt);
}
// one warp; returns the number of reset envs in the tiles in front of tile gt (uniform over the warp)
__device__ __noinline__ unsigned lookback_resolve(unsigned long long* st, const unsigned epoch, const int gt, const unsigned cnt, const int lane) {
  const unsigned long long tag = (unsigned long long)epoch << 32;
  if (gt == 0) return 0u;
  unsigned excl = 0;
  int idx = gt - 1;
  constexpr int kWin = 4;   // windows of 32 tiles read per round: one L2 round trip covers the 128 tiles of a 4096-env launch
  for (;;) {
    unsigned long long w[kWin];
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const int j = idx - 32 * k - lane;   // lane 0 of window 0 looks at the nearest tile
      w[k] = (j >= 0) ? ld_relaxed_u64(st + j) : (tag | (2ull << 24));   // in front of tile 0: prefix 0
    }
    bool done = false;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      if (done) break;
      const int j = idx - 32 * k - lane;
      int tries = 0;
      for (;;) {
        const bool ok = ((unsigned)(w[k] >> 32) == epoch) && ((((unsigned)w[k]) >> 24) & 3u) != 0u;
        if (__all_sync(0xffffffffu, ok)) break;
        if (++tries > (1 << 22)) __trap();   // a predecessor never published: fail loudly instead of hanging
        w[k] = (j >= 0) ? ld_relaxed_u64(st + j) : (tag | (2ull << 24));
      }
      const unsigned is_p = __ballot_sync(0xffffffffu, ((((unsigned)w[k]) >> 24) & 3u) == 2u);
      unsigned v = ((unsigned)w[k]) & 0xffffffu;
      if (is_p != 0u && lane > __ffs(is_p) - 1) v = 0u;   // nothing beyond the nearest inclusive prefix
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      excl += v;
      done = (is_p != 0u);
    }
    if (done) break;
    idx -= 32 * kWin;
  }
  if (lane == 0) st_relaxed_u64(st + gt, tag | (2ull << 24) | (excl + cnt));
  return excl;
}

// ---------------------------------------------------------------------------------------------------
// PRE: TerminationManager.compute + RewardManager.compute [IL] + reset_buf.nonzero()
// ---------------------------------------------------------------------------------------------------
// resident CTAs per SM the register allocation aims at (8 / 16 warps per CTA)
#ifndef RL_V2_PRE_MINB8
#define RL_V2_PRE_MINB8 5
#endif
#ifndef RL_V2_PRE_MINB16
#define RL_V2_PRE_MINB16 2
#endif
#ifndef RL_V2_POST_MINB8
#define RL_V2_POST_MINB8 6
#endif
#ifndef RL_V2_POST_MINB16
#define RL_V2_POST_MINB16 3
#endif
template <class B, int C, int G, int NW>
__global__ void __launch_bounds__(NW * 32, G > 2 ? 1 : (NW <= 4 ? 8 : (NW <= 8 ? RL_V2_PRE_MINB8 : RL_V2_PRE_MINB16))) v2_pre_kernel(const __grid_constant__ V2Args a) {
  using CF = Cfg2<B, RL_V2_PRE, C, G, NW>;
  constexpr int kWarps2 = NW, kThreads2 = NW * 32;
  constexpr Layout L = CF::L;
  constexpr Scalars S = CF::S;
  constexpr int E = CF::E, W = CF::W, K = S.num_reward_terms;
  constexpr bool CN = true;   // HIST_MAX_NORM reads the cached norms
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = warp / W, slot = warp - tile * W;
  const int e = tile * 32 + lane;
  const uint32_t role = C > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = (int)blockIdx.x / C;
  const int env0 = cluster_id * E;
  const long long env = (long long)env0 + e;
  const bool my_hist = ((CF::hist_roles >> role) & 1u) != 0 && S.num_hist_bodies > 0;
#if RL_V2_STAMPS
  if (a.k.dbg == reinterpret_cast<long long*>(1)) return;   // launch-overhead probe (tools/v2_timeline.py --empty)
#endif
  V2_GTIME(0); V2_STAMP_T0(1);
  if (C > 1) cluster_arrive_relaxed();   // "every CTA of the cluster runs": waited for in front of the first DSMEM store
  if (a.k.use_pdl) pdl_launch_dependents();   // the successor's prologue may overlap this kernel
  if (tid == 0) init_load_barrier(&s_bar);
  V2_STAMP_T0(9);
  __syncthreads();                            // the armed mbarrier is visible to every issuing warp
  V2_STAMP_T0(10);
  if (a.k.use_pdl) pdl_wait();                // no global access of this kernel before its predecessor is complete
  if (lane == 0) {
    issue_loads<CF>(sm, a, role, env0, &s_bar, warp);
    // the post-reset launch of this env step streams the tile's ray hits: have them in L2 by then
    if (role == 0 && warp == kWarps2 - 1 && a.prefetch_rays != nullptr)
      prefetch_l2(a.prefetch_rays + (size_t)env0 * a.prefetch_row_bytes, (uint32_t)(E * a.prefetch_row_bytes));
  }
  V2_STAMP_T0(11);
  if (my_hist)   // contact-force norms from global memory, in the shadow of the record's copies
    norm_prepass<CF>(sm, static_cast<const float*>(a.k.hist.ptr) + (size_t)env0 * (S.hist_len * S.num_hist_bodies * 3), warp, lane);
  // per-joint constants: device table -> shared
  for (int i = tid; i < 5 * S.num_joints; i += kThreads2) sm[L.cj + i] = __ldg(a.k.cj + i);   // [5][J] table, one coalesced read
  V2_STAMP_T0(12);
  if (role == 0)   // weight-0 terms: no task writes their slot of the final sum
    for (int i = tid; i < K * E; i += kThreads2)
      if ((a.k.rw_zero >> (i / E)) & 1ull) sm[L.termv + i] = 0.f;
  __syncthreads();           // mbarrier init + the stores above visible to the CTA
  V2_STAMP_T0(2);
  mbar_wait(&s_bar, 0);      // record resident
#if RL_V2_STAMPS
  if (a.k.dbg == reinterpret_cast<long long*>(2)) return;   // probe: launch + load only
#endif
  V2_STAMP_T0(3);
  if (C > 1) cluster_wait_acquire();
  V2_STAMP_T0(4);

  float* const termv0 = C > 1 ? map_to_rank(sm + L.termv, 0u) : sm + L.termv;   // role 0's term values
  const FieldD f_sums = a.k.outf[OF_SUMS], f_stepr = a.k.outf[OF_STEPR];
  // In a cluster the weighted value goes to role 0 first (DSMEM), the CTA's arrival at the cluster barrier follows, and the
  // global result rows leave AFTER it: an arrive.release waits for the thread's earlier stores, global ones included.
  constexpr int kPend = 6;
  float p_sum[kPend], p_step[kPend];
  int p_k[kPend], np = 0;
  unsigned lb_mask = 0, lb_epoch = 0;   // the termination task's warp: done mask of the tile, epoch of the look-back ...
  int lb_gt = -1;                       // ... and the tile's index (-1: not that warp)
  dispatch_bin<B, CF, 0, CF::BINS>((int)role * W + slot, sm, e,
      [&](auto, const Task& tk, const RlRewardTerm& rt, const RlObsTerm&, const bool, const EnvCtx& c) __attribute__((always_inline)) {
    if (tk.kind == TK_REWARD) {
      const int k = tk.a;
      const float sum_old = __ldg(static_cast<const float*>(f_sums.ptr) + (size_t)k * f_sums.cs + env);   // in flight while the term is evaluated
      const float raw = reward_term<CN>(rt, c_spec[a.k.slot].rewards[tk.a], S, L, sm, e, c, 0, 64);
      // RewardManager.compute [IL]: value = func * weight * dt; sums += value; step_reward = value / dt
      const float val = (raw * rt.weight) * S.step_dt;
      termv0[k * E + e] = val;
      const float ns = sum_old + val, sr = rl_div(val, S.step_dt);
      if (C > 1 && np < kPend) {
        p_sum[np] = ns; p_step[np] = sr; p_k[np] = k; ++np;
      } else {
        static_cast<float*>(const_cast<void*>(f_sums.ptr))[(size_t)k * f_sums.cs + env] = ns;
        if (f_stepr.ptr) static_cast<float*>(const_cast<void*>(f_stepr.ptr))[(size_t)k * f_stepr.cs + env] = sr;
      }
    } else if (tk.kind == TK_DONES) {
      // TerminationManager.compute [IL]: bits | terminated << 8 | time_out << 9
      const unsigned scan_epoch = ld_relaxed_u32(a.scan_ctl);   // in flight while the terms are evaluated
      const int eplen_now = __float_as_int(SMF(L.eplen, 0)) + 1;
      static_cast<int*>(const_cast<void*>(a.k.outf[OF_EPLEN].ptr))[env] = eplen_now;
      uint32_t bits = 0, term = 0, trunc = 0;
      StaticPolicy<B>::for_dones(a.k, [&](const RlDoneTerm& t, int d) __attribute__((always_inline)) {
        int fired = 0;
        if (t.type == RL_DONE_TIME_OUT) {
          fired = eplen_now >= S.max_episode_length;
        } else if (t.type == RL_DONE_TERRAIN_OUT_OF_BOUNDS) {
          fired = (t.p[2] != 0.f) && ((fabsf(c.pos.x) > t.p[0]) || (fabsf(c.pos.y) > t.p[1]));
        } else if (t.type == RL_DONE_ILLEGAL_CONTACT) {
          RL_FEW_LOOP
          for (int b = 0; b < S.num_hist_bodies; ++b)
            if (((t.body_mask >> b) & 1ull) && (SMF(L.hnorm, b) > t.p[0])) fired = 1;
        }
        if (fired) { bits |= 1u << d; if (t.time_out) trunc = 1; else term = 1; }
      });
      SMF(L.flags, 0) = __int_as_float((int)(bits | (term << 8) | (trunc << 9)));
      if (a.k.out.done_bits) a.k.out.done_bits[env] = (uint8_t)bits;
      if (a.k.out.terminated) a.k.out.terminated[env] = (uint8_t)term;
      if (a.k.out.truncated) a.k.out.truncated[env] = (uint8_t)trunc;
      // reset_buf.nonzero() [IL]: this tile's ids at its offset into the ascending list (decoupled look-back, no tail)
      lb_mask = __ballot_sync(0xffffffffu, (term | trunc) != 0);
      lb_gt = cluster_id * G + tile;
      lb_epoch = scan_epoch;
      lookback_publish(a.scan_state, scan_epoch, lb_gt, __popc(lb_mask), lane);   // resolved after the task barrier
    }
  }, [&]() __attribute__((always_inline)) {
    if (C > 1) cluster_arrive_release();   // this warp's term values are in role 0's record
#pragma unroll
    for (int i = 0; i < kPend; ++i)
      if (i < np) {
        static_cast<float*>(const_cast<void*>(f_sums.ptr))[(size_t)p_k[i] * f_sums.cs + env] = p_sum[i];
        if (f_stepr.ptr) static_cast<float*>(const_cast<void*>(f_stepr.ptr))[(size_t)p_k[i] * f_stepr.cs + env] = p_step[i];
      }
  }, V2_DBG_ROW);
  V2_STAMP(16 + warp);
  if (C > 1) { if (role != 0) return; cluster_wait_acquire(); }
  __syncthreads();   // the term values of role 0's own warps, the termination flags
  V2_STAMP_T0(5);

  // ---- final sum: one warp per tile adds the weighted values up in manager order (is_terminated is finished here) ----
  if (slot == 0) {
    const int fl = __float_as_int(SMF(L.flags, 0));
    if (CF::late != 0) {
#pragma unroll 1
      for (uint64_t m = CF::late; m != 0; m &= m - 1) {
        const int k = __ffsll((long long)m) - 1;
        const float raw = ((fl >> 8) & 1) ? 1.f : 0.f;
        const float val = (raw * a.k.rw_weight[k]) * S.step_dt;
        SMF(L.termv, k) = val;
        float* gsum = static_cast<float*>(const_cast<void*>(f_sums.ptr)) + (size_t)k * f_sums.cs + env;
        *gsum = *gsum + val;
        if (f_stepr.ptr) static_cast<float*>(const_cast<void*>(f_stepr.ptr))[(size_t)k * f_stepr.cs + env] = rl_div(val, S.step_dt);
      }
    }
    float tv[K > 0 ? K : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) tv[k] = SMF(L.termv, k);    // all K loads in flight, then the serial sum
    float total = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) total += tv[k];             // manager order
    if (a.k.out.reward) a.k.out.reward[env] = total;
    if (f_stepr.ptr)
#pragma unroll 1
      for (uint64_t m = a.k.rw_zero; m != 0; m &= m - 1)
        static_cast<float*>(const_cast<void*>(f_stepr.ptr))[(size_t)(__ffsll((long long)m) - 1) * f_stepr.cs + env] = 0.f;
  }
  // ---- reset_buf.nonzero() [IL]: the tile's ids at its offset into the ascending list (decoupled look-back, no tail) ----
  if (lb_gt >= 0) {   // warp-uniform: the termination task's warp (not the final sum's when the tile has more than one warp)
    const unsigned cnt = __popc(lb_mask);
    const unsigned excl = lookback_resolve(a.scan_state, lb_epoch, lb_gt, cnt, lane);
    if (((lb_mask >> lane) & 1u) && a.k.out.reset_ids) a.k.out.reset_ids[excl + __popc(lb_mask & ((1u << lane) - 1u))] = (int32_t)env;
    if (lb_gt == a.k.vgrid - 1 && lane == 0) {   // the last tile: every tile has published (and read the epoch) by now
      if (a.k.out.n_reset) *a.k.out.n_reset = (int32_t)(excl + cnt);
      a.scan_ctl[0] = lb_epoch + 1u;
    }
  }
  V2_STAMP_T0(6); V2_GTIME(7);
}

// ---------------------------------------------------------------------------------------------------
// POST: ManagerBasedRLEnv._reset_idx (manager part) for the done envs + CommandManager.compute +
// ObservationManager.compute [IL] for all envs
// ---------------------------------------------------------------------------------------------------
template <class B, int C, int G, int NW>
__global__ void __launch_bounds__(NW * 32, G > 2 ? 1 : (NW <= 4 ? 8 : (NW <= 8 ? 6 : 3))) v2_post_kernel(const __grid_constant__ V2Args a) {
  using CF = Cfg2<B, RL_V2_POST, C, G, NW>;
  constexpr int kWarps2 = NW, kThreads2 = NW * 32;
  constexpr Layout L = CF::L;
  constexpr Scalars S = CF::S;
  constexpr int E = CF::E, W = CF::W, K = S.num_reward_terms, A = S.n_actions;
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = warp / W, slot = warp - tile * W;
  const int e = tile * 32 + lane;
  const uint32_t role = C > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = (int)blockIdx.x / C;
  const int env0 = cluster_id * E;
  const long long env = (long long)env0 + e;
#if RL_V2_STAMPS
  if (a.k.dbg == reinterpret_cast<long long*>(1)) return;
#endif
  V2_GTIME(0); V2_STAMP_T0(1);
  if (C > 1) cluster_arrive_relaxed();
  if (a.k.use_pdl) pdl_launch_dependents();
  for (int i = tid; i < 5 * S.num_joints; i += kThreads2) sm[L.cj + i] = __ldg(a.k.cj + i);   // [5][J] table, one coalesced read
  if (tid == 0) {
    s_last = 0;
    init_load_barrier(&s_bar);
  }
  __syncthreads();
  if (a.k.use_pdl) pdl_wait();
  if (lane == 0) issue_loads<CF>(sm, a, role, env0, &s_bar, warp);
  // byte flags of this lane's env (every role needs the reset mask; role 0 also the command flags)
  const int u8_reset = (a.k.out.terminated[env] | a.k.out.truncated[env]) != 0;
  int u8_head = 0, u8_stand = 0;
  if (role == 0 && slot == 0) {
    u8_head = static_cast<const uint8_t*>(a.k.is_heading.ptr)[env];
    u8_stand = static_cast<const uint8_t*>(a.k.is_standing.ptr)[env];
  }
  const int u8_bits = (a.k.out.done_bits != nullptr) ? (int)a.k.out.done_bits[env] : 0;   // LOG tasks; requested early
  RandState rs;
  rs.seed = a.k.rnd.seed;
  rs.step = a.k.rnd.step + (a.k.rnd.step_counter ? *a.k.rnd.step_counter : 0ull);
  rs.env_id_offset = a.k.rnd.env_id_offset;

  // ---- height scan: global -> registers -> global, lanes = columns ------------------------------------------------
  // height_scan [IL] = sensor z - ray hit z - offset, then clip, then scale (no noise: v2_spec_ok). One CTA per tile: right
  // here, while the record is still in flight. In a cluster: after the cluster barrier (a release arrival would wait for
  // these ~190 stores per row).
  auto stream_height_scan = [&]() __attribute__((always_inline)) {
  static_for(std::make_integer_sequence<int, RL_NUM_OBS_GROUPS>{}, [&](auto gc) {
    constexpr int g = decltype(gc)::value;
    constexpr int st = scan_term_of(B::spec.obs[g]);
    if constexpr (st >= 0) {
      static constexpr RlObsTerm t = B::spec.obs[g].terms[st];
      constexpr int col0 = obs_col0(B::spec.obs[g], st);
      constexpr int R = t.dim;
      if (a.k.out.obs[g] != nullptr) {
        constexpr int rows_per_cta = E / C;   // this role's share of the cluster's env rows
        const float* rays = static_cast<const float*>(a.k.rays.ptr);
        const float* rz = static_cast<const float*>(a.k.in[IF_RAYPOS].ptr);
        constexpr int RPW = (rows_per_cta + kWarps2 - 1) / kWarps2;   // rows per warp
        constexpr int CH = (R + 31) / 32;                              // 32-column chunks per row
        float z[RPW], v[RPW][CH];
#pragma unroll
        for (int i = 0; i < RPW; ++i) {   // every load of the warp's rows is in flight before the first use
          const int r = warp + i * kWarps2;
          const long long ev = (long long)env0 + (int)role * rows_per_cta + r;
          if (r < rows_per_cta) {
            z[i] = __ldg(rz + ev);
#pragma unroll
            for (int c = 0; c < CH; ++c) v[i][c] = (c * 32 + lane < R) ? __ldg(rays + ev * R + c * 32 + lane) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const int r = warp + i * kWarps2;
          const long long ev = (long long)env0 + (int)role * rows_per_cta + r;
          if (r < rows_per_cta) {
            float* dst = a.k.out.obs[g] + ev * a.k.out.obs_pitch[g] + col0;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const int col = c * 32 + lane;
              if (col < R) {
                float x = (z[i] - v[i][c]) - t.p[0];
                if (t.has_clip) x = clampf(x, t.clip_lo, t.clip_hi);
                if (t.has_scale) x = x * t.scale;
                dst[col] = x;
              }
            }
          }
        }
      }
    }
  });
  };
  if (C == 1) stream_height_scan();

  if (slot == 0 && role == 0) {   // the command task's warp: its own lanes read these back
    sm[L.ishead + e] = __int_as_float(u8_head); sm[L.isstand + e] = __int_as_float(u8_stand);
  }
  __syncthreads();
  V2_STAMP_T0(2);
  mbar_wait(&s_bar, 0);
#if RL_V2_STAMPS
  if (a.k.dbg == reinterpret_cast<long long*>(2)) return;
#endif
  V2_STAMP_T0(3);

  // ---- tasks: no barrier between the load and the row stores ------------------------------------------------------
  // The manager reset of the envs flagged done (RewardManager / ActionManager / CommandTerm .reset [IL]) is part of the
  // tasks that own the state it touches: the LOG task reduces the logging partials of its tile and zeroes the reset envs'
  // episode sums / stored actions / episode length in global memory, the COMMAND task resamples their command before its
  // own update, the observation tasks see a reset env's stored action as 0 and its episode length as 0.
  const bool rme = u8_reset != 0;
  unsigned early_prev = 0xffffffffu;
  const int n_tiles = a.k.N / 32;
  V2_STAMP_T0(4);
  const int eplen_now = rme ? 0 : __float_as_int(SMF(L.eplen, 0));
  if (C > 1) cluster_wait_acquire();   // every CTA of the cluster runs: DSMEM may be written
  bool arrived = false;                // this thread's arrival at the second cluster barrier (as early as its DSMEM stores allow)
  dispatch_bin<B, CF, 0, CF::BINS>((int)role * W + slot, sm, e,
      [&](auto, const Task& tk, const RlRewardTerm&, const RlObsTerm& ot, const bool corrupt, const EnvCtx& c) __attribute__((always_inline)) {
    if (tk.kind == TK_OBS) {
      if (a.k.out.obs[tk.a] == nullptr) return;
      obs_task(sm, L, S, ot, c_spec[a.k.slot].obs[tk.a].terms[tk.b], corrupt, a.k, rs, tk.a, tk.b, tk.col0, tk.lo, tk.hi, e, env, c, eplen_now,
               nullptr, rme);
    } else if (tk.kind == TK_LOG) {
      // logging partials of the tile (combined by the last tile to arrive, in a fixed order -> deterministic):
      // every quantity is reduced over the lanes (= envs) with a fixed shuffle tree
      if (C > 1 && !arrived) { cluster_arrive_release(); arrived = true; }   // nothing of this task goes through DSMEM
      const int gt = cluster_id * G + tile;
      const int part = tk.b, parts = tk.col0;
      const int fl = u8_bits;
      constexpr int NQ = K + RL_MAX_DONE_TERMS + 2;
      if (__ballot_sync(0xffffffffu, rme) == 0u) {   // nothing to reset in this tile: the partials are zero
        for (int q = part + lane * parts; q < NQ; q += 32 * parts) a.k.log_partials[(size_t)gt * RL_LOG_STRIDE + q] = 0.f;
      } else {
#pragma unroll 4
        for (int q = part; q < NQ; q += parts) {   // independent reductions: their shuffle trees overlap
          float x = 0.f;
          if (rme) {
            if (q < K) x = sm[L.sums + q * E + e];
            else if (q < K + RL_MAX_DONE_TERMS) x = (float)((fl >> (q - K)) & 1);
            else x = sm[(q == K + RL_MAX_DONE_TERMS ? L.mxy : L.myaw) + e];
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
          if (lane == 0) a.k.log_partials[(size_t)gt * RL_LOG_STRIDE + q] = x;
        }
      }
      __syncwarp();
      if (lane == 0) early_prev = ticket_arrive_release(a.k.ticket);   // the tile's partials are written (the release covers only them)
      // ... and only now the zeroing of the reset envs' rows: the ticket's release does not have to wait for these stores
      if (rme) {
        float* gs = static_cast<float*>(const_cast<void*>(a.k.outf[OF_SUMS].ptr));
        float* ga = static_cast<float*>(const_cast<void*>(a.k.outf[OF_ACT].ptr));
        float* gp = static_cast<float*>(const_cast<void*>(a.k.outf[OF_PACT].ptr));
#pragma unroll 1
        for (int k = part; k < K; k += parts) gs[(size_t)k * a.k.outf[OF_SUMS].cs + env] = 0.f;
#pragma unroll 1
        for (int q = part; q < A; q += parts) {
          ga[(size_t)q * a.k.outf[OF_ACT].cs + env] = 0.f;
          gp[(size_t)q * a.k.outf[OF_PACT].cs + env] = 0.f;
        }
        if (part == 0) static_cast<int*>(const_cast<void*>(a.k.outf[OF_EPLEN].ptr))[env] = 0;
      }
    } else if (tk.kind == TK_COMMAND) {
      // CommandTerm.reset [IL] of the reset envs (resample: V/mdp/commands.py:43-47), then CommandManager.compute and the
      // observation columns that show the new command (written into their owners' rows). The metric accumulators work on
      // a private copy (the LOG task reads the old values concurrently).
      constexpr Layout LC = [] { Layout l = CF::L; l.mxy = CF::L.rmask; l.myaw = CF::L.epnew; return l; }();
      sm[LC.mxy + e] = rme ? 0.f : sm[L.mxy + e];
      sm[LC.myaw + e] = rme ? 0.f : sm[L.myaw + e];
      EnvCtx cc = c;
      if (rme) {
        constexpr RlCommandCfg cfgc = B::spec.command;
        float u[RL_NUM_CMD_UNIFORMS];
        if (a.k.rnd.cmd_uniforms != nullptr) {
#pragma unroll
          for (int q = 0; q < RL_NUM_CMD_UNIFORMS; ++q) u[q] = sm[L.cmdu + q * E + e];
        } else {
          const uint4 r0 = rl_philox(rs, env, RL_STREAM_RESET_COMMAND, 0), r1 = rl_philox(rs, env, RL_STREAM_RESET_COMMAND, 1);
          u[0] = u01(r0.x); u[1] = u01(r0.y); u[2] = u01(r0.z); u[3] = u01(r0.w);
          u[4] = u01(r1.x); u[5] = u01(r1.y); u[6] = u01(r1.z);
        }
        float c0 = u[1] * (cfgc.lin_vel_x_hi - cfgc.lin_vel_x_lo) + cfgc.lin_vel_x_lo;
        float c1 = u[2] * (cfgc.lin_vel_y_hi - cfgc.lin_vel_y_lo) + cfgc.lin_vel_y_lo;
        const float c2 = u[3] * (cfgc.ang_vel_z_hi - cfgc.ang_vel_z_lo) + cfgc.ang_vel_z_lo;
        const float keep = (sqrtf(c0 * c0 + c1 * c1) > cfgc.small_cmd_threshold) ? 1.f : 0.f;
        c0 *= keep; c1 *= keep;
        cc.c0 = c0; cc.c1 = c1; cc.c2 = c2;
        sm[L.tleft + e] = u[0] * (cfgc.resampling_time_hi - cfgc.resampling_time_lo) + cfgc.resampling_time_lo;
        if (cfgc.heading_command) {
          sm[L.head + e] = u[4] * (cfgc.heading_hi - cfgc.heading_lo) + cfgc.heading_lo;
          sm[L.ishead + e] = __int_as_float((u[5] <= cfgc.rel_heading_envs) ? 1 : 0);
        }
        sm[L.isstand + e] = __int_as_float((u[6] <= cfgc.rel_standing_envs) ? 1 : 0);
      }
      command_update(sm, LC, S, StaticPolicy<B>::command(a.k), a.k, rs, e, env, cc, true);
      StaticPolicy<B>::for_cmd_obs(a.k, [&](const RlObsTerm& t, int g, int ti, int col0, bool corr) __attribute__((always_inline)) {
        if (a.k.out.obs[g] == nullptr) return;
        const uint32_t owner = (uint32_t)(g == 0 ? CF::obs_owner0 : CF::obs_owner1);
        float* smo = (C > 1 && owner != 0) ? map_to_rank(sm, owner) : sm;   // the owner's record: same offsets
        obs_task(sm, LC, S, t, c_spec[a.k.slot].obs[g].terms[ti], corr, a.k, rs, g, ti, col0, 0, t.dim, e, env, cc, eplen_now, smo);
      });
      if (C > 1 && !arrived) { cluster_arrive_release(); arrived = true; }   // the command columns are in their owners' rows
      // command state of the step (CommandTerm fields), straight from registers / the record
      float* gc = static_cast<float*>(const_cast<void*>(a.k.outf[OF_CMD].ptr));
      gc[env] = SMF(LC.cmdn, 0); gc[(size_t)a.k.outf[OF_CMD].cs + env] = SMF(LC.cmdn, 1); gc[2 * (size_t)a.k.outf[OF_CMD].cs + env] = SMF(LC.cmdn, 2);
      static_cast<float*>(const_cast<void*>(a.k.outf[OF_HEAD].ptr))[env] = sm[L.head + e];
      static_cast<float*>(const_cast<void*>(a.k.outf[OF_TLEFT].ptr))[env] = sm[L.tleft + e];
      static_cast<float*>(const_cast<void*>(a.k.outf[OF_MXY].ptr))[env] = sm[LC.mxy + e];
      static_cast<float*>(const_cast<void*>(a.k.outf[OF_MYAW].ptr))[env] = sm[LC.myaw + e];
      static_cast<uint8_t*>(const_cast<void*>(a.k.is_heading.ptr))[env] = (uint8_t)__float_as_int(sm[L.ishead + e]);
      static_cast<uint8_t*>(const_cast<void*>(a.k.is_standing.ptr))[env] = (uint8_t)__float_as_int(sm[L.isstand + e]);
    }
  }, [&]() __attribute__((always_inline)) {
    if (C > 1 && !arrived) cluster_arrive_release();
  }, V2_DBG_ROW);
  V2_STAMP(16 + warp);
  if (lane == 0 && early_prev == (unsigned)(n_tiles * CF::log_parts - 1)) { ticket_acquire(a.k.ticket); s_last = 1; }   // the last ticket of the launch
  if (C > 1) cluster_wait_acquire();   // the command columns have reached their rows
  __syncthreads();

  V2_STAMP_T0(5);
  // ---- observation rows: this role's groups, per-env columns, lanes = columns (coalesced) -------------------------
  static_for(std::make_integer_sequence<int, RL_NUM_OBS_GROUPS>{}, [&](auto gc) {
    constexpr int g = decltype(gc)::value;
    constexpr int D = env_cols_of(B::spec.obs[g]);
    if constexpr (D > 0) {
      if ((uint32_t)(g == 0 ? CF::obs_owner0 : CF::obs_owner1) == role && a.k.out.obs[g] != nullptr) {
        const float* rows = sm + LOBS(g);
        constexpr int P = LOBSP(g);
#pragma unroll 1
        for (int r = warp; r < E; r += kWarps2) {
          float* dst = a.k.out.obs[g] + ((long long)env0 + r) * a.k.out.obs_pitch[g];
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 32)
            if (c0 + lane < D) dst[c0 + lane] = rows[r * P + c0 + lane];
        }
      }
    }
  });

  if (C > 1) stream_height_scan();
  V2_STAMP_T0(6); V2_GTIME(7);
  // ---- logging means of the reset (extras["log"] [IL]): the CTA whose ticket was the last of the launch -------------
  if (s_last) {   // CTA-uniform (any role: the CTA whose LOG warp drew the last ticket; its thread acquired the tiles' rows)
    float* s_red = sm;   // the record is dead
    __syncthreads();
    for (int q = lane; q < K + RL_MAX_DONE_TERMS + 2; q += 32)
      for (int vw = warp; vw < kLogWarps; vw += kWarps2) {   // 16 strided partial sums whatever the warp count of this CTA
        float part = 0.f;
        for (int g = vw; g < n_tiles; g += kLogWarps) part += __ldcg(a.k.log_partials + (size_t)g * RL_LOG_STRIDE + q);
        s_red[vw * RL_LOG_STRIDE + q] = part;
      }
    __syncthreads();
    if (tid < K + RL_MAX_DONE_TERMS + 2) {
      float tot = 0.f;
      for (int w = 0; w < kLogWarps; ++w) tot += s_red[w * RL_LOG_STRIDE + tid];
      const RlResetLog& lg = a.k.out.reset_log;
      const int n_reset_total = *a.k.out.n_reset;
      const float cnt = (float)max(n_reset_total, 1);
      if (n_reset_total == 0) tot = 0.f;
      if (tid < K) { if (lg.episode_sum_mean) lg.episode_sum_mean[tid] = tot / cnt; }
      else if (tid < K + RL_MAX_DONE_TERMS) { if (lg.done_term_count) lg.done_term_count[tid - K] = tot; }
      else if (lg.metric_mean) lg.metric_mean[tid - K - RL_MAX_DONE_TERMS] = tot / cnt;
    }
    if (tid == 0) *a.k.ticket = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side: tensor maps, eligibility, launch
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct TmKey {
  const void* ptr;
  int cs, nc, n, box;
};
struct TmEntry {
  TmKey key;
  CUtensorMap map;
};
constexpr int kTmCache = 96;   // per field: the bench rotates over 24 state sets, tests over a handful

}  // namespace

struct RlV2State {
  int baked;
  long long launches;        // launches the cluster kernels handled (rl_ctx_get_cluster_config)
  int force_c, force_g, force_nw;   // RL_MDPSTEP_V2_CFG="CxGxNW" pins a configuration (A/B measurements); 0 = choose by env count
  int large_n, large_nw;            // from large_n envs on: large_nw warps per tile
  int n_tm[IF_COUNT];
  int next_tm[IF_COUNT];
  TmEntry tm[IF_COUNT][kTmCache];
  // device scratch of the look-back (V2Args::scan_state / scan_ctl)
  unsigned long long* scan_state;
  int scan_cap;
  unsigned int* scan_ctl;
};

namespace {

// tensor map of one SoA field: [nc, N] fp32, row stride cs elements, box {box envs, nc components}
int tensor_map_for(RlV2State* v, int f, const FieldD& fd, int nc, int n, int box, CUtensorMap* out) {
  const TmKey key{fd.ptr, fd.cs, nc, n, box};
  for (int i = 0; i < v->n_tm[f]; ++i) {
    const TmKey& k = v->tm[f][i].key;
    if (k.ptr == key.ptr && k.cs == key.cs && k.nc == key.nc && k.n == key.n && k.box == key.box) { *out = v->tm[f][i].map; return RL_OK; }
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(RL_ECUDA, "cuTensorMapEncodeTiled is not available through cudaGetDriverEntryPoint%s", "");
  // single-component fields: the (unused) row stride still has to be a multiple of 16 bytes
  const unsigned long long row_bytes = nc > 1 ? (unsigned long long)fd.cs * 4ull : (((unsigned long long)n * 4ull + 15ull) & ~15ull);
  const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)nc};
  const cuuint64_t gstr[1] = {(cuuint64_t)row_bytes};
  const cuuint32_t bdim[2] = {(cuuint32_t)box, (cuuint32_t)nc};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(fd.ptr), gdim, gstr, bdim, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RL_ECUDA, "cuTensorMapEncodeTiled failed for input field %s%lld (CUresult %lld)", "", f, (long long)r);
  int slot = v->n_tm[f] < kTmCache ? v->n_tm[f]++ : (v->next_tm[f]++ % kTmCache);
  v->tm[f][slot].key = key; v->tm[f][slot].map = m;
  *out = m;
  return RL_OK;
}

bool soa_ok(const FieldD& d, int nc) {   // what a tensor map (or a lane = env access) can describe
  return d.ptr != nullptr && d.es == 1 && (reinterpret_cast<uintptr_t>(d.ptr) & 15u) == 0 && (nc == 1 || (d.cs % 4) == 0);
}

template <class B, int KIND, int C, int G, int NW>
int launch_v2(RlCtx* ctx, const KArgs& k, cudaStream_t st) {
  using CF = Cfg2<B, KIND, C, G, NW>;
  RlV2State* v = ctx->v2;
  const RlStepSpec& s = ctx->spec;
  V2Args a;
  memset(&a, 0, sizeof(a));
  a.k = k;
  a.k.vgrid = k.N / 32;
  const uint32_t staged = v2_field_mask(s, KIND) & k.in_mask;
  for (int f = 0; f < IF_COUNT; ++f) {
    if (!((staged >> f) & 1u)) continue;
    int rc = tensor_map_for(v, f, k.in[f], in_field_ncomp(s, f), k.N, CF::E, &a.tm[f]);
    if (rc != RL_OK) return rc;
    a.field_word[f] = v2_field_word(s, KIND, f);
  }
  for (int r = 0; r < C; ++r) {
    const uint32_t m = role_field_mask(s, CF::sched, KIND, r, CF::W) & staged;
    uint32_t bytes = 0;
    int n = 0;
    for (int f = 0; f < IF_COUNT; ++f)
      if ((m >> f) & 1u) { bytes += (uint32_t)(in_field_ncomp(s, f) * CF::E * 4); a.role_field[r][n++] = (uint8_t)f; }
    // largest copies first: they are the last to land
    std::stable_sort(a.role_field[r], a.role_field[r] + n, [&](uint8_t x, uint8_t y) { return in_field_ncomp(s, x) > in_field_ncomp(s, y); });
    a.role_n[r] = (uint8_t)n; a.role_bytes[r] = bytes;
  }
  a.scan_state = v->scan_state; a.scan_ctl = v->scan_ctl;
  if (KIND == RL_V2_PRE && s.num_rays > 0 && k.rays.ptr != nullptr && k.rays.cs == 1 && k.rays.es == s.num_rays &&
      ((size_t)CF::E * s.num_rays * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(k.rays.ptr) & 15u) == 0) {
    a.prefetch_rays = static_cast<const char*>(k.rays.ptr);
    a.prefetch_row_bytes = (uint32_t)s.num_rays * 4u;
  }
  auto kern = KIND == RL_V2_PRE ? v2_pre_kernel<B, C, G, NW> : v2_post_kernel<B, C, G, NW>;
  const size_t smem = (size_t)CF::L.total_words * 4;
  static thread_local int configured_device = -1;
  if (configured_device != ctx->device) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_device = ctx->device;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(k.N / CF::E) * C); cfg.blockDim = dim3(CF::NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = k.use_pdl ? 2 : 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
  return RL_OK;
}

// configurations compiled for every baked spec: (cluster size, tiles per CTA, warps per CTA)
#define RL_V2_CONFIGS(X) X(1, 1, 16) X(1, 1, 8) X(1, 1, 4) X(2, 2, 16) X(4, 4, 16)

template <class B, int KIND>
int dispatch_v2_cfg(RlCtx* ctx, const KArgs& k, int c, int g, int nw, cudaStream_t st, bool* found) {
  *found = true;
#define RL_CASE(C_, G_, NW_) if (c == (C_) && g == (G_) && nw == (NW_)) return launch_v2<B, KIND, C_, G_, NW_>(ctx, k, st);
  RL_V2_CONFIGS(RL_CASE)
#undef RL_CASE
  *found = false;
  return RL_OK;
}

bool cfg_compiled(int c, int g, int nw) {
#define RL_CASE(C_, G_, NW_) if (c == (C_) && g == (G_) && nw == (NW_)) return true;
  RL_V2_CONFIGS(RL_CASE)
#undef RL_CASE
  return false;
}

// The configuration a launch of n envs uses (measured on B200, profiles/r2_summary.md): up to a few waves of tiles one
// tile per CTA with 16 warps (shortest serial chain per tile); beyond that fewer warps per tile, so that more tiles are
// in flight per SM and less per-warp context work is repeated.
void choose_cfg(const RlV2State* v, int64_t n, int* c, int* g, int* nw) {
  *c = 0; *g = 0; *nw = 0;
  if (n <= 0 || n % 32 != 0) return;
  if (v->force_c > 0) {
    if (n % (32 * v->force_g) == 0) { *c = v->force_c; *g = v->force_g; *nw = v->force_nw; }
    return;
  }
  *c = 1; *g = 1;
  *nw = n >= v->large_n ? v->large_nw : 16;
}

#if RL_V2_DEV_ONE
#define RL_V2_BAKED_LIST(X) X(Baked6)
#else
#define RL_V2_BAKED_LIST(X) RL_BAKED_LIST(X)
#endif

}  // namespace

int rl_v2_create(RlCtx* ctx) {
  ctx->v2 = nullptr;
  if (ctx->baked < 0) return RL_OK;
  const char* sw = getenv("RL_MDPSTEP_V2");
  if (sw && sw[0] == '0') return RL_OK;
  if (!v2_spec_ok(ctx->spec)) return RL_OK;
#if RL_V2_DEV_ONE
  if (memcmp(&ctx->spec, &baked::Baked6::spec, sizeof(RlStepSpec)) != 0) return RL_OK;
#endif
  RlV2State* v = new (std::nothrow) RlV2State();
  if (!v) return fail(RL_ENOMEM, "rl_ctx_create: out of host memory%s", "");
  memset(v, 0, sizeof(*v));
  v->baked = ctx->baked;
  v->large_n = 16384; v->large_nw = 8;
  const char* cfg = getenv("RL_MDPSTEP_V2_CFG");   // "CxG" (16 warps) or "CxGxNW"
  if (cfg && cfg[0]) {
    int c = 0, g = 0, nw = 16;
    const int got = sscanf(cfg, "%dx%dx%d", &c, &g, &nw);
    if (got < 2 || !cfg_compiled(c, g, nw)) {
      delete v;
      return fail(RL_EINVAL, "RL_MDPSTEP_V2_CFG=%s is not a compiled (cluster size x tiles per CTA [x warps]) configuration", cfg);
    }
    v->force_c = c; v->force_g = g; v->force_nw = nw;
  }
  // this translation unit's copy of the spec slots (the term functions read run-time indexed lists from it)
  CUDA_TRY(cudaMemcpyToSymbol(c_spec, &ctx->spec, sizeof(RlStepSpec), sizeof(RlStepSpec) * ctx->slot));
  ctx->v2 = v;   // from here on rl_v2_destroy frees what has been allocated
  CUDA_TRY(cudaMalloc(&v->scan_ctl, sizeof(unsigned int) * 64));
  CUDA_TRY(cudaMemset(v->scan_ctl, 0, sizeof(unsigned int) * 64));
  const unsigned int first_epoch = 1u;   // zeroed status words are never valid
  CUDA_TRY(cudaMemcpy(v->scan_ctl, &first_epoch, sizeof(first_epoch), cudaMemcpyHostToDevice));
  return RL_OK;
}

void rl_v2_destroy(RlCtx* ctx) {
  if (ctx->v2) {
    if (ctx->v2->scan_state) cudaFree(ctx->v2->scan_state);
    if (ctx->v2->scan_ctl) cudaFree(ctx->v2->scan_ctl);
  }
  delete ctx->v2;
  ctx->v2 = nullptr;
}

namespace {
// status words of the look-back: zeroed once (epoch tags make the words of earlier launches invalid, never a stale word of
// uninitialised memory); grows with the env count
int ensure_scan_state(RlV2State* v, int n_tiles) {
  if (n_tiles <= v->scan_cap) return RL_OK;
  if (v->scan_state) CUDA_TRY(cudaFree(v->scan_state));
  v->scan_state = nullptr; v->scan_cap = 0;
  const int cap = n_tiles * 2 + 64;
  CUDA_TRY(cudaMalloc(&v->scan_state, sizeof(unsigned long long) * (size_t)cap));
  CUDA_TRY(cudaMemset(v->scan_state, 0, sizeof(unsigned long long) * (size_t)cap));
  v->scan_cap = cap;
  return RL_OK;
}
}  // namespace

void rl_v2_config_for(const RlCtx* ctx, int64_t num_envs, int* cluster_size, int* tiles_per_cta, int* warps_per_cta, long long* launches) {
  *cluster_size = 0; *tiles_per_cta = 0; *warps_per_cta = 0; *launches = 0;
  if (!ctx->v2) return;
  choose_cfg(ctx->v2, num_envs, cluster_size, tiles_per_cta, warps_per_cta);
  *launches = ctx->v2->launches;
}

int rl_v2_try_launch(RlCtx* ctx, const KArgs& k, int kind, cudaStream_t st, bool* handled) {
  *handled = false;
  RlV2State* v = ctx->v2;
  if (!v || k.has_ids || (k.dbg != nullptr && !RL_V2_STAMPS)) return RL_OK;
  const RlStepSpec& s = ctx->spec;
  int c = 0, g = 0, nw = 0;
  choose_cfg(v, k.N, &c, &g, &nw);
  if (c == 0) return RL_OK;
  // every staged field must be an SoA tensor a tensor map can describe; the sensor rows one contiguous block
  const uint32_t want = v2_field_mask(s, kind);
  for (int f = 0; f < IF_COUNT; ++f) {
    if (!((want >> f) & 1u)) continue;
    if (f == IF_CMDU && k.in[f].ptr == nullptr) continue;   // production mode: in-kernel Philox
    if (!((k.in_mask >> f) & 1u) || !soa_ok(k.in[f], in_field_ncomp(s, f))) return RL_OK;
  }
  if (kind == RL_V2_PRE) {
    const int HW = s.hist_len * s.num_hist_bodies * 3;
    if (HW > 0 && (k.hist.ptr == nullptr || k.hist.cs != 1 || k.hist.es != HW)) return RL_OK;   // contiguous force rows
    if (!soa_ok(k.outf[OF_SUMS], s.num_reward_terms) || !k.outf[OF_EPLEN].ptr || k.outf[OF_EPLEN].es != 1) return RL_OK;
    if (k.outf[OF_STEPR].ptr && k.outf[OF_STEPR].es != 1) return RL_OK;
  } else {
    if (!k.out.terminated || !k.out.truncated || !k.out.n_reset) return RL_OK;
    if (!k.is_heading.ptr || !k.is_standing.ptr || k.is_heading.es != 1 || k.is_standing.es != 1) return RL_OK;
    for (int f : {OF_SUMS, OF_ACT, OF_PACT, OF_CMD, OF_HEAD, OF_TLEFT, OF_MXY, OF_MYAW, OF_EPLEN})
      if (!k.outf[f].ptr || k.outf[f].es != 1) return RL_OK;
    if (s.num_rays > 0) {
      bool scan = false;
      for (int gi = 0; gi < RL_NUM_OBS_GROUPS; ++gi) scan = scan || (scan_term_of(s.obs[gi]) >= 0 && k.out.obs[gi] != nullptr);
      if (scan && (k.rays.ptr == nullptr || k.rays.cs != 1 || k.rays.es != s.num_rays || k.in[IF_RAYPOS].ptr == nullptr || k.in[IF_RAYPOS].es != 1))
        return RL_OK;
    }
  }
  if (k.N / 32 >= (1 << 24)) return RL_OK;   // the look-back's status word holds a 24-bit count
  if (kind == RL_V2_PRE) {
    const int rcs = ensure_scan_state(v, k.N / 32);
    if (rcs != RL_OK) return rcs;
  }
  bool found = false;
  int rc = RL_OK, idx = 0;
#if RL_V2_DEV_ONE
  idx = 6;
#endif
#define RL_TRY(B_)                                                                                        \
  if (idx++ == v->baked) {                                                                               \
    rc = kind == RL_V2_PRE ? dispatch_v2_cfg<baked::B_, RL_V2_PRE>(ctx, k, c, g, nw, st, &found)             \
                           : dispatch_v2_cfg<baked::B_, RL_V2_POST>(ctx, k, c, g, nw, st, &found);           \
  }
  RL_V2_BAKED_LIST(RL_TRY)
#undef RL_TRY
  *handled = found;
  if (found && rc == RL_OK) ++v->launches;
  return rc;
}
