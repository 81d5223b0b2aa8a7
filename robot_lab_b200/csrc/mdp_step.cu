// mdp_step.cu - fused per-step MDP pipeline for sm_100a (B200).
//
// One launch evaluates, for every env of a CTA's tile: termination terms, all reward terms (+ episode sums
// and per-term step rewards), the velocity-command update, both observation groups and the ordered
// compaction of reset ids. Reference behaviour restated (paths relative to /root/reference, V/ =
// source/robot_lab/robot_lab/tasks/manager_based/locomotion/velocity/): V/mdp/rewards.py:22-687,
// V/mdp/observations.py:17-35, V/mdp/commands.py:22-85, V/velocity_env_cfg.py:106-254,379-664 and the
// IsaacLab manager loops / upstream terms listed in SURVEY.md Appendix A.
//
// Execution model (HBM-bound by arithmetic intensity, ~0.5 FLOP/B, no tensor cores; at a few thousand envs bound
// by first-touch latency - see DESIGN.md 2 and profiles/r1_summary.md):
//   * a CTA owns a tile of 32 consecutive envs; in the compute phase lane e of every warp IS env e (no lane
//     repeats another lane's per-env scalar work) and the warps differ in which terms they evaluate (a constexpr
//     longest-processing-time schedule, reached through one binary-search branch per warp);
//   * load phase: the AoS sensor rows (contact-force history, height-scan ray hits) are staged into shared
//     memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), the small per-env SoA fields with
//     16-byte cp.async into an SoA [word][32] record - all global reads of the tile are in flight before any
//     arithmetic starts, one join;
//   * stage 1 works on shared memory only, every warp finishes its terms (weight, dt, episode sum, step reward)
//     and assembles its observation columns in the shared-memory rows; stage 2 (warp 0) adds the reward up in
//     manager order;
//   * store phase: observation rows leave with bulk stores (cp.async.bulk.global.shared::cta), the SoA outputs
//     with 16-byte stores; the last CTA to finish compacts the reset ids from per-CTA bit masks / combines the
//     reset-logging partials.
//
// Built with -fmad=false on purpose: the reference is eager PyTorch, every op rounds on its own, and not
// contracting a*b+c keeps threshold decisions (contact > 1 N, |cmd| > 0.1, ...) bit-identical.

#include "rl_common.cuh"

#include <new>
#include <utility>

#include "generated/baked_specs.cuh"

#define RL_SPEC_SLOTS 3
#define RL_PI_F 3.14159265358979323846f
#define RL_LOG_STRIDE 64  // >= RL_MAX_REWARD_TERMS + RL_MAX_DONE_TERMS + 2

__constant__ RlStepSpec c_spec[RL_SPEC_SLOTS];

namespace {

#define fail rl_fail
#define g_err g_rl_err
using DeviceGuard = RlDeviceGuard;

// ---------------------------------------------------------------------------------------------------
// Per-launch field descriptors
//
// A "row" is one component of one 4-byte per-env field. The step kernel stages rows into an SoA shared-memory
// record (word w of local env e at sm[w*32 + e]) with non-blocking cp.async copies and writes result rows back the
// same way. Each of the ~25 fields travels per launch as {pointer, strides, component count, first record word}
// in the kernel parameter bank - one parameter line per field, no table in global memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kE = 32;  // envs per CTA = lanes per warp: in the compute phase lane e of every warp owns env e
// A reward term is evaluated in at most this many parts (termv holds that many slots per term). 2, not more: finer
// cuts measured slower, and every extra slot costs K * 128 bytes of the tile record - at 4 the Go2-rough record grew
// from 114.6 to 119.9 KB and lost the second resident CTA per SM (1.5 x slower from 16 k envs up)
// RL_SHARED_NORMS=1 (build variant, robot_lab_b200/build.py --variant shared_norms; NOT the default, not yet measured on a
// GPU): the max-over-history contact-force norm of every body is computed ONCE per env by a prepass in which all warps
// run the same code (body b by warp b mod NW, lane = env) and kept in the record (hnorm); undesired_contacts,
// contact_forces, feet_slide and illegal_contact read it instead of each recomputing T norms per body inside its own
// single-warp chain. Same function, same operand order: bit-identical values. With the long body sums gone no
// term is split, so termv needs one slot per term and the record does not grow.
#ifndef RL_SHARED_NORMS
#define RL_SHARED_NORMS 0
#endif
#if RL_SHARED_NORMS
constexpr int kTermParts = 1;
#else
constexpr int kTermParts = 2;
#endif
// RL_SHARED_CTX=1 (build variant shared_ctx, also part of variant "shared"; not the default, not yet measured): the three
// quaternion rotations every warp needs for its lane's env (projected gravity, base-frame linear / angular velocity)
// are computed once by warps 0-2 and published in the record (ctxv) instead of sixteen times; same functions, same
// operands: bit-identical values. One barrier in front of stage 1 (shared with the norm prepass when both are on).
#ifndef RL_SHARED_CTX
#define RL_SHARED_CTX 0
#endif
// RL_PERSISTENT=1 (build variant persistent; not the default, not yet measured): a launch never has more CTAs than the
// GPU holds at once; a CTA walks tiles blockIdx, blockIdx + gridDim, ... with the same code, so from its second tile on
// the kernel's instructions are already in the SM's caches - the multi-wave regime (> ~4700 envs), where every wave of
// fresh CTAs starts cold today (profiles/r1_summary.md section 5). One tile per CTA per iteration, record reused.
#ifndef RL_PERSISTENT
#define RL_PERSISTENT 0
#endif

struct FieldD {
  const void* ptr;
  int es;    // env stride   (elements)
  int cs;    // comp stride  (elements)
  int meta;  // movers: components | first record word << 16 (kept next to the pointer: one parameter-bank line)
  int pad_;
};

enum InField {
  IF_ROOT_POS = 0, IF_QUAT, IF_LIN_VEL, IF_ANG_VEL, IF_JPOS, IF_JVEL, IF_JACC, IF_JTAU,
  IF_CAIR, IF_LAIR, IF_CCON, IF_LCON, IF_BPOS, IF_BVEL, IF_RAYPOS,
  IF_CMD, IF_HEAD, IF_TLEFT, IF_MXY, IF_MYAW, IF_EPLEN, IF_SUMS, IF_CMDU, IF_ACT, IF_PACT, IF_COUNT
};
enum OutField { OF_REWARD = 0, OF_EPLEN, OF_SUMS, OF_STEPR, OF_CMD, OF_HEAD, OF_TLEFT, OF_MXY, OF_MYAW, OF_ACT, OF_PACT, OF_COUNT };


// ---------------------------------------------------------------------------------------------------
// Shared-memory layout of one CTA tile (word offsets; SoA words already multiplied by kE).
// ---------------------------------------------------------------------------------------------------
struct Layout {
  int A, J, K;
  int root_pos, quat, lin_vel, ang_vel;          // SoA offsets (= word * kE)
  int jpos, jvel, jacc, jtau;
  int act, pact;
  int cmd, head, tleft, ishead, isstand;
  int rmask;                                      // 1 = this env is being reset by the launch (RESET phase)
  int cmdn, epnew;                                // updated command / episode length (committed by the store phase)
  int mxy, myaw, eplen;
  int sums;
  int cair, lair, ccon, lcon;
  int bpos, bvel;
  int raypos;
  int cmdu;
  int rew, flags, stepr;                         // outputs
  int termv;                                     // [K][kTermParts] weighted term value in slot 0 (raw partial sums of a split term before it is finished)
  int arrive;                                    // [K] per-env arrival counters of the two halves of a split term
#if RL_SHARED_NORMS
  int hnorm;                                     // [B] max over the history of |F_b| (written by the prepass)
#endif
#if RL_SHARED_CTX
  int ctxv;                                      // [9] projected gravity, base-frame lin vel, base-frame ang vel
#endif
  int w_rew, w_eplen, w_sums, w_stepr, w_cmd, w_head, w_tleft, w_mxy, w_myaw, w_act, w_pact;  // record words of out fields
  int soa_words;
  int cj;                                        // per-joint constants [5][J]: q0, qd0, soft lo, soft hi, vel limit
  // AoS rows [kE][pitch]; pitches are forced ODD so that lane e reading row e is bank-conflict free
  int hist, hist_pitch;
  int rays, rays_pitch;
  int obs0, obs1, obs_pitch0, obs_pitch1;          // selected with LOBS(g) etc.: no runtime-indexed members, so
                                                   // the struct never has to live in local memory
  int total_words;
};

__host__ __device__ constexpr int in_field_ncomp(const RlStepSpec& s, int f) {
  switch (f) {
    case IF_ROOT_POS: case IF_LIN_VEL: case IF_ANG_VEL: case IF_CMD: return 3;
    case IF_QUAT: return 4;
    case IF_JPOS: case IF_JVEL: case IF_JACC: case IF_JTAU: return s.num_joints;
    case IF_CAIR: case IF_LAIR: case IF_CCON: case IF_LCON: return s.num_time_bodies;
    case IF_BPOS: case IF_BVEL: return 3 * s.num_asset_bodies;
    case IF_SUMS: return s.num_reward_terms;
    case IF_CMDU: return RL_NUM_CMD_UNIFORMS;
    case IF_ACT: case IF_PACT: return s.action.n_actions;
    default: return 1;
  }
}

__host__ __device__ constexpr int in_field_word(const RlStepSpec& s, int f) {  // first record word of an input field
  int w = 0;
  for (int i = 0; i < f; ++i) w += in_field_ncomp(s, i);
  return w;
}
__host__ __device__ constexpr int out_field_ncomp(const RlStepSpec& s, int f) {
  switch (f) {
    case OF_SUMS: case OF_STEPR: return s.num_reward_terms;
    case OF_CMD: return 3;
    case OF_ACT: case OF_PACT: return s.action.n_actions;
    default: return 1;
  }
}

__host__ __device__ constexpr int align_up(int v, int a) { return (v + a - 1) / a * a; }
static_assert(kE == 32, "the field movers assume 8 float4 per row");
__host__ __device__ constexpr int odd_pitch(int n) { return n <= 0 ? 1 : (n | 1); }

__host__ __device__ constexpr Layout make_layout(const RlStepSpec& s) {
  Layout L{};
  constexpr int E = kE;
  int w = 0;
  auto take = [&w](int n) { int o = w; w += n; return o; };
  const int J = s.num_joints, A = s.action.n_actions, K = s.num_reward_terms;
  L.A = A; L.J = J; L.K = K;
  int in_word[IF_COUNT] = {};
  for (int f = 0; f < IF_COUNT; ++f) in_word[f] = take(in_field_ncomp(s, f));
  L.root_pos = in_word[IF_ROOT_POS] * E; L.quat = in_word[IF_QUAT] * E;
  L.lin_vel = in_word[IF_LIN_VEL] * E; L.ang_vel = in_word[IF_ANG_VEL] * E;
  L.jpos = in_word[IF_JPOS] * E; L.jvel = in_word[IF_JVEL] * E; L.jacc = in_word[IF_JACC] * E; L.jtau = in_word[IF_JTAU] * E;
  L.cair = in_word[IF_CAIR] * E; L.lair = in_word[IF_LAIR] * E; L.ccon = in_word[IF_CCON] * E; L.lcon = in_word[IF_LCON] * E;
  L.bpos = in_word[IF_BPOS] * E; L.bvel = in_word[IF_BVEL] * E; L.raypos = in_word[IF_RAYPOS] * E;
  L.cmd = in_word[IF_CMD] * E; L.head = in_word[IF_HEAD] * E; L.tleft = in_word[IF_TLEFT] * E;
  L.mxy = in_word[IF_MXY] * E; L.myaw = in_word[IF_MYAW] * E; L.eplen = in_word[IF_EPLEN] * E;
  L.sums = in_word[IF_SUMS] * E; L.cmdu = in_word[IF_CMDU] * E;
  L.act = in_word[IF_ACT] * E; L.pact = in_word[IF_PACT] * E;
  L.w_sums = in_word[IF_SUMS];
  L.w_head = in_word[IF_HEAD]; L.w_tleft = in_word[IF_TLEFT]; L.w_mxy = in_word[IF_MXY]; L.w_myaw = in_word[IF_MYAW];
  L.w_act = in_word[IF_ACT]; L.w_pact = in_word[IF_PACT];
  L.ishead = take(1) * E; L.isstand = take(1) * E;
  L.rmask = take(1) * E;
  // results the step commits at the end: tasks of the same stage still read the old command / episode length
  { const int wn = take(3); L.cmdn = wn * E; L.w_cmd = wn; }
  { const int wn = take(1); L.epnew = wn * E; L.w_eplen = wn; }
  L.w_rew = take(1); L.rew = L.w_rew * E;
  L.flags = take(1) * E;
  L.w_stepr = take(K); L.stepr = L.w_stepr * E;
  L.termv = take(kTermParts * K) * E;
  L.arrive = take(K) * E;
#if RL_SHARED_NORMS
  L.hnorm = take(s.num_hist_bodies) * E;
#endif
#if RL_SHARED_CTX
  L.ctxv = take(9) * E;
#endif
  L.soa_words = w;
  int off = align_up(w * E, 32);  // 128-byte aligned sections (bulk copies need 16 B)
  L.cj = off; off = align_up(off + 5 * J, 32);
  L.hist_pitch = odd_pitch(s.hist_len * s.num_hist_bodies * 3);
  L.hist = off; off = align_up(off + E * L.hist_pitch, 32);
  L.rays_pitch = odd_pitch(s.num_rays);
  L.rays = off; off = align_up(off + E * L.rays_pitch, 32);
  L.obs_pitch0 = odd_pitch(s.obs[0].dim); L.obs_pitch1 = odd_pitch(s.obs[1].dim);
  L.obs0 = off; off = align_up(off + E * L.obs_pitch0, 32);
  L.obs1 = off; off = align_up(off + E * L.obs_pitch1, 32);
  L.total_words = off;
  return L;
}

__host__ __device__ constexpr int out_field_word(const Layout& L, int f) {
  switch (f) {
    case OF_REWARD: return L.w_rew; case OF_EPLEN: return L.w_eplen; case OF_SUMS: return L.w_sums;
    case OF_STEPR: return L.w_stepr; case OF_CMD: return L.w_cmd; case OF_HEAD: return L.w_head;
    case OF_TLEFT: return L.w_tleft; case OF_MXY: return L.w_mxy; case OF_MYAW: return L.w_myaw;
    case OF_ACT: return L.w_act; default: return L.w_pact;
  }
}

// ---------------------------------------------------------------------------------------------------
// Work schedule. The compute phase is thread-per-env (lane e of every warp owns env e, so SIMT lanes never
// duplicate per-env scalar work); the warps of a CTA differ in WHICH tasks they run: one task per reward term
// (wide body-mask terms split in two halves) and one per observation term (the height scan in 64-column chunks),
// balanced over the warps by a longest-processing-time greedy on rough instruction costs. constexpr, so a baked
// spec gets its schedule at compile time and every warp's code is straight-line.
// ---------------------------------------------------------------------------------------------------
enum { TK_REWARD = 0, TK_OBS = 1, TK_DONES = 2, TK_COMMAND = 3 };

struct Task {
  uint8_t kind, a, b, owner;   // REWARD: a = term k, b = half (0/1); OBS: a = group, b = term index
  uint16_t lo, hi;             // REWARD: body-index range [lo, hi); OBS: column range within the term
  uint16_t col0, pad;          // OBS: first column of the term inside the group row; REWARD: col0 = number of parts,
                               // pad = 1 for one part of a split term
};
struct Schedule {
  int n;
  Task t[RL_MAX_TASKS];
  uint8_t split[RL_MAX_REWARD_TERMS];   // > 0: the term is evaluated in that many parts (termv[k][0..parts))
  uint8_t late[RL_MAX_REWARD_TERMS];    // term is finished in stage 2 (split terms, is_terminated)
};

__host__ __device__ constexpr int popc64(uint64_t m) { int n = 0; while (m) { m &= m - 1; ++n; } return n; }

__host__ __device__ constexpr int reward_cost(const RlRewardTerm& t, const RlStepSpec& s, int nbodies) {
  const int J = popc64(t.joint_mask), F = t.n_idx, T = s.hist_len;
  switch (t.type) {
    case RL_REW_JOINT_TORQUES_L2: case RL_REW_JOINT_VEL_L2: case RL_REW_JOINT_ACC_L2: case RL_REW_JOINT_DEVIATION_L1:
    case RL_REW_JOINT_POWER: case RL_REW_STAND_STILL: return 30 + 5 * J;
    case RL_REW_JOINT_POS_LIMITS: case RL_REW_JOINT_VEL_LIMITS: case RL_REW_JOINT_POS_PENALTY: return 40 + 8 * J;
    case RL_REW_JOINT_MIRROR: case RL_REW_ACTION_MIRROR: return 30 + 8 * F;
    case RL_REW_ACTION_SYNC: return 40 + 30 * F;
    case RL_REW_ACTION_RATE_L2: return 30 + 5 * s.action.n_actions;
#if RL_SHARED_NORMS
    case RL_REW_UNDESIRED_CONTACTS: case RL_REW_CONTACT_FORCES: return 30 + nbodies * 5;
#else
    case RL_REW_UNDESIRED_CONTACTS: case RL_REW_CONTACT_FORCES: return 30 + nbodies * T * 18;
#endif
    case RL_REW_TRACK_LIN_VEL_XY_EXP: case RL_REW_TRACK_ANG_VEL_Z_EXP: case RL_REW_TRACK_ANG_VEL_Z_WORLD_EXP: return 70;
    case RL_REW_TRACK_LIN_VEL_XY_YAW_FRAME_EXP: return 260;
    case RL_REW_FEET_AIR_TIME: case RL_REW_FEET_CONTACT: case RL_REW_FEET_CONTACT_WITHOUT_CMD: return 30 + 10 * F;
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: return 40 + 12 * F;
    case RL_REW_FEET_AIR_TIME_VARIANCE: return 40 + 40 * F;
    case RL_REW_FEET_GAIT: return 260;
    case RL_REW_FEET_STUMBLE: return 30 + 20 * F;
#if RL_SHARED_NORMS
    case RL_REW_FEET_SLIDE: return 30 + F * 62;
#else
    case RL_REW_FEET_SLIDE: return 30 + F * (60 + T * 18);
#endif
    case RL_REW_FEET_HEIGHT: return 40 + 60 * F;
    case RL_REW_FEET_HEIGHT_BODY: return 40 + 130 * F;
    case RL_REW_FEET_DISTANCE_Y_EXP: case RL_REW_FEET_DISTANCE_XY_EXP: return 80 + 60 * F;
    case RL_REW_WHEEL_VEL_PENALTY: return 40 + 12 * F;
    default: return 30;
  }
}

__host__ __device__ constexpr Schedule make_schedule(const RlStepSpec& s, int nw) {
  Schedule sc{};
  int cost[RL_MAX_TASKS] = {};
  int n = 0;
  sc.t[n] = Task{TK_DONES, 0, 0, 0, 0, 0, 0, 0}; cost[n++] = 90;
  {
    int c = 380;  // command update + heading control (+ the termination terms again when done envs are skipped)
    for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g)
      for (int ti = 0; ti < s.obs[g].n_terms; ++ti)
        if (s.obs[g].terms[ti].type == RL_OBS_GENERATED_COMMANDS) c += 40;   // it also owns these columns
    sc.t[n] = Task{TK_COMMAND, 0, 0, 0, 0, 0, 0, 0}; cost[n++] = c;
  }
  for (int k = 0; k < s.num_reward_terms; ++k) {
    const RlRewardTerm& t = s.rewards[k];
    if (t.weight == 0.f) continue;
    if (t.type == RL_REW_IS_TERMINATED) { sc.late[k] = 1; continue; }
    // Long sums over bodies / feet can be cut into up to kTermParts parts; the parts publish raw partial sums and
    // the one that arrives last (per env) adds them in part order. Measured (profiles/r1_summary.md): every part is
    // its own straight-line code and costs ~4k cycles of first-touch instruction fetch whatever its length, so
    // only the one really long term (> 8 bodies) is cut, and only in two.
    const bool body_sum = (t.type == RL_REW_UNDESIRED_CONTACTS || t.type == RL_REW_CONTACT_FORCES);
    const bool list_sum = (t.type == RL_REW_FEET_SLIDE);
    const int nb = popc64(t.body_mask);
    const int items = body_sum ? nb : (list_sum ? t.n_idx : 0);
    int parts = items > 8 ? 2 : 1;
    if (parts > kTermParts) parts = kTermParts;
    if (n + parts > RL_MAX_TASKS - 24) parts = 1;   // table nearly full: stop splitting
    if (parts < 2) {
      sc.t[n] = Task{TK_REWARD, (uint8_t)k, 0, 0, 0, 64, 1, 0}; cost[n++] = 30 + reward_cost(t, s, body_sum ? nb : t.n_idx);
      continue;
    }
    sc.split[k] = (uint8_t)parts;
    int item0 = 0;
    for (int p = 0; p < parts; ++p) {
      const int item1 = (items * (p + 1)) / parts;   // items [item0, item1) belong to part p
      int lo = item0, hi = item1;
      if (body_sum) {          // body terms take a body-index range: translate item counts into bit positions
        int seen = 0; lo = 64; hi = 64;
        for (int bb = 0; bb < 64; ++bb)
          if ((t.body_mask >> bb) & 1ull) {
            if (seen == item0) lo = bb;
            if (seen == item1) hi = bb;
            ++seen;
          }
        if (p == 0) lo = 0;
        if (p == parts - 1) hi = 64;
      } else if (p == parts - 1) {
        hi = 64;
      }
      sc.t[n] = Task{TK_REWARD, (uint8_t)k, (uint8_t)p, 0, (uint16_t)lo, (uint16_t)hi, (uint16_t)parts, 1};
      cost[n++] = reward_cost(t, s, item1 - item0);
      item0 = item1;
    }
  }
  for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
    int col0 = 0;
    for (int ti = 0; ti < s.obs[g].n_terms; ++ti) {
      const RlObsTerm& o = s.obs[g].terms[ti];
      const int per_col = 8 + ((o.has_noise && s.obs[g].enable_corruption) ? 24 : 0);
      if (o.type != RL_OBS_GENERATED_COMMANDS) {
        // multiples of 4 (one Philox block = 4 columns); coarser when the task table is nearly full
        const int chunk = (n > RL_MAX_TASKS - 24) ? 256 : 64;
        for (int lo = 0; lo < o.dim; lo += chunk) {
          const int hi = (lo + chunk < o.dim) ? lo + chunk : o.dim;
          sc.t[n] = Task{TK_OBS, (uint8_t)g, (uint8_t)ti, 0, (uint16_t)lo, (uint16_t)hi, (uint16_t)col0, 0};
          cost[n++] = 20 + per_col * (hi - lo);
        }
      }
      col0 += o.dim;
    }
  }
  sc.n = n;
  // longest-processing-time greedy over the warps, separately for the two task classes an env step runs in
  // separate launches (terminations + rewards before the reset, command + observations after it): each launch
  // sees a balanced schedule, and so does a launch that runs everything
  for (int cls = 0; cls < 2; ++cls) {
    int load[32] = {};
    bool done[RL_MAX_TASKS] = {};
    for (int it = 0; it < n; ++it) {
      int best = -1;
      for (int i = 0; i < n; ++i) {
        const bool in_cls = ((sc.t[i].kind == TK_REWARD || sc.t[i].kind == TK_DONES) ? 0 : 1) == cls;
        if (in_cls && !done[i] && (best < 0 || cost[i] > cost[best])) best = i;
      }
      if (best < 0) break;
      // ties go to the higher warp for the second class so that a launch running both does not pile up on warp 0
      int w = cls == 0 ? 0 : nw - 1;
      for (int j = 0; j < nw; ++j) {
        const int jj = cls == 0 ? j : nw - 1 - j;
        if (load[jj] < load[w]) w = jj;
      }
      sc.t[best].owner = (uint8_t)w; load[w] += cost[best]; done[best] = true;
    }
  }
  return sc;
}

struct KArgs {
  int N;
  int slot;
  uint32_t phases;
  int has_ids;
  FieldD in[IF_COUNT];
  FieldD outf[OF_COUNT];
  uint32_t in_mask, out_mask;
  uint32_t in_vec4, out_vec4;     // fields whose rows may move 4 envs at a time (SoA, 16-byte aligned)
  float rw_weight[RL_MAX_REWARD_TERMS];          // stage 2: weights and term classes come from the parameter bank
  uint64_t rw_late, rw_isterm, rw_zero;   // bit k: finished in stage 2 / is_terminated / weight 0
  // AoS spans (row-contiguous per env) and byte fields
  FieldD hist, rays;
  FieldD is_heading, is_standing;               // uint8
  RlStepOut out;
  RlRandom rnd;
  const int32_t* env_ids;
  const int32_t* n_env_ids;
  Layout L;                // generic kernel only; baked kernels compute theirs at compile time
  const Schedule* sched;   // device copy (generic kernel); baked kernels carry theirs as constexpr data
  int vgrid;               // number of tiles ("virtual CTAs") of the launch; == gridDim.x when a CTA carries one tile
  int tile_words;          // distance between the records of a CTA's tiles in shared memory (words; multi-tile CTAs)
  unsigned int* ticket;
  uint32_t* cta_mask;
  float* log_partials;   // [grid][RL_LOG_STRIDE] per-CTA partial sums of the reset logging reductions
  int use_pdl;
  long long* dbg;        // optional [grid][RL_DEBUG_STRIDE] clock64 stamps (rl_ctx_set_debug_buffer)
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
// Ampere-style async copies (SASS LDGSTS): global -> shared without a register round trip, fire and forget
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// "last CTA" tickets: the increment is a release at device scope (everything this thread wrote or observed through a
// CTA barrier before it is visible to whoever reads the final count and then fences) - one MEMBAR.ALL.GPU instead of
// the sequentially-consistent fence of __threadfence() (MEMBAR.SC.GPU + L1 invalidation), and no result is needed
// until the tail, so the issuing warp does not wait for the round trip
__device__ __forceinline__ unsigned ticket_arrive_release(unsigned int* ticket) {
  unsigned prev;
  asm volatile("atom.add.release.gpu.u32 %0, [%1], 1;" : "=r"(prev) : "l"(ticket) : "memory");
  return prev;
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
__device__ __noinline__ float rl_div(float x, float y) { return x / y; }   // IEEE division, one code copy
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
  int obs_dim0, obs_dim1;
};
__host__ __device__ constexpr Scalars scalars_of(const RlStepSpec& s) {
  return Scalars{s.num_joints, s.num_hist_bodies, s.hist_len, s.num_time_bodies, s.num_asset_bodies, s.num_rays,
                 s.num_reward_terms, s.num_done_terms, s.max_episode_length, s.action.n_actions,
                 s.step_dt, s.contact_time_abs_tol, s.obs[0].dim, s.obs[1].dim};
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
  __device__ __forceinline__ static Layout layout(const KArgs& a) { return a.L; }
  __device__ __forceinline__ static Scalars scalars(const KArgs& a) { return scalars_of(c_spec[a.slot]); }
  __device__ __forceinline__ static const RlCommandCfg& command(const KArgs& a) { return c_spec[a.slot].command; }
  // f(tag, task, reward term, obs term, corruption on, task index) for every task the warp owns
  template <int NW, class F> __device__ __forceinline__ static void for_tasks(const KArgs& a, int warp, F&& f) {
    const RlStepSpec& S = c_spec[a.slot];
    const int n = a.sched->n;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const Task tk = a.sched->t[i];
      if (tk.owner != warp) continue;
      if (tk.kind == TK_OBS) f(std::integral_constant<int, -1>{}, tk, S.rewards[0], S.obs[tk.a].terms[tk.b], S.obs[tk.a].enable_corruption != 0, i);
      else f(std::integral_constant<int, -2>{}, tk, S.rewards[tk.a], S.obs[0].terms[0], false, i);
    }
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
  template <class F> __device__ __forceinline__ static void for_dones(const KArgs& a, F&& f) {
    const RlStepSpec& S = c_spec[a.slot];
#pragma unroll 1
    for (int d = 0; d < S.num_done_terms; ++d) f(S.dones[d], d);
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
  using Baked = B;
  __device__ __forceinline__ static constexpr Layout layout(const KArgs&) {
    constexpr Layout L = make_layout(B::spec);
    return L;
  }
  __device__ __forceinline__ static constexpr Scalars scalars(const KArgs&) {
    constexpr Scalars s = scalars_of(B::spec);
    return s;
  }
  __device__ __forceinline__ static constexpr RlCommandCfg command(const KArgs&) {
    constexpr RlCommandCfg c = B::spec.command;
    return c;
  }
  template <int NW> struct Sched { static constexpr Schedule value = make_schedule(B::spec, NW); };
  // all tasks of warp W, in schedule order; one lambda instantiation (= one call site) per task
  template <int NW, int W, class F> __device__ __forceinline__ static void warp_tasks(F&& f) {
    static_for(std::make_integer_sequence<int, Sched<NW>::value.n>{}, [&](auto ic) {
      constexpr int i = decltype(ic)::value;
      constexpr Task tk = Sched<NW>::value.t[i];
      if constexpr (tk.owner == W) {
        static constexpr RlRewardTerm rt = B::spec.rewards[tk.kind == TK_REWARD ? tk.a : 0];
        static constexpr RlObsTerm ot = B::spec.obs[tk.kind == TK_OBS ? tk.a : 0].terms[tk.kind == TK_OBS ? tk.b : 0];
        constexpr bool corrupt = B::spec.obs[tk.kind == TK_OBS ? tk.a : 0].enable_corruption != 0;
        f(ic, tk, rt, ot, corrupt, i);
      }
    });
  }
  // binary search on the warp id: log2(NW) uniform branches lead to the warp's own contiguous code, so a warp
  // never walks (or fetches) the code of the others
  template <int NW, int LO, int HI, class F> __device__ __forceinline__ static void dispatch_warp(int warp, F&& f) {
    if constexpr (HI - LO == 1) {
      warp_tasks<NW, LO>(f);
    } else {
      constexpr int MID = (LO + HI) / 2;
      if (warp < MID) dispatch_warp<NW, LO, MID>(warp, f); else dispatch_warp<NW, MID, HI>(warp, f);
    }
  }
  template <int NW, class F> __device__ __forceinline__ static void for_tasks(const KArgs&, int warp, F&& f) {
    dispatch_warp<NW, 0, NW>(warp, f);
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
  template <class F> __device__ __forceinline__ static void for_dones(const KArgs&, F&& f) {
    static_for(std::make_integer_sequence<int, B::spec.num_done_terms>{}, [&](auto dc) {
      constexpr int d = decltype(dc)::value;
      static constexpr RlDoneTerm t = B::spec.dones[d];
      f(t, d);
    });
  }
  template <class F> __device__ __forceinline__ static void for_joint_consts(const KArgs& a, int j, F&& f) {
    DynPolicy::for_joint_consts(a, j, f);  // lane-varying joint index: staged from __constant__ like the generic path
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

#define LOBS(g) ((g) == 0 ? L.obs0 : L.obs1)
#define LOBSP(g) ((g) == 0 ? L.obs_pitch0 : L.obs_pitch1)
#define SMF(off, c) sm[(off) + (c) * kE + e]
#define CJ(k, j) sm[L.cj + (k) * L.J + (j)]

__device__ __forceinline__ bool first_contact(const float* sm, const Layout& L, const Scalars& S, int e, int b) {
  const float t = SMF(L.ccon, b);
  return (t > 0.f) && (t < (S.step_dt + S.contact_time_abs_tol));
}
__device__ __forceinline__ V3 body_vec(const float* sm, int off, int e, int b) {
  return V3{SMF(off, 3 * b + 0), SMF(off, 3 * b + 1), SMF(off, 3 * b + 2)};
}
// what the norm consumers call: the record's cached value (prepass) or the computation itself
#if RL_SHARED_NORMS
#define HIST_MAX_NORM(h, b) SMF(L.hnorm, (b))
#else
#define HIST_MAX_NORM(h, b) hist_max_norm(h, S.hist_len, S.num_hist_bodies, b)   /* arguments are plain identifiers / subscripts */
#endif
// max over the history of |F_b| (net_forces_w_history[:, :, b].norm(-1).max(1))
__device__ __noinline__ float hist_max_norm(const float* h, int T, int B, int b) {
  // ONE copy for every term that uses it (undesired_contacts, contact_forces, feet_slide, feet_stumble, the
  // illegal-contact termination): warps in different terms keep the same few instruction lines hot instead of
  // evicting each other's private copies. A force of exactly 0 skips the IEEE sqrt (its special-case path).
  float m = 0.f;
  _Pragma("unroll 1")
  for (int t = 0; t < T; ++t) {
    const float* f = h + (t * B + b) * 3;
    const float ss = (f[0] * f[0] + f[1] * f[1]) + f[2] * f[2];
    const float n = (ss == 0.f) ? 0.f : sqrtf(ss);
    m = (t == 0) ? n : fmaxf(m, n);
  }
  return m;
}

// One reward term for env e: raw value (no weight, no dt). [lo, hi) restricts body-mask terms to a body-index
// range (the two halves of a split term add up).
// `t` may be a build-time constant (scalar members fold into immediates); `tc` is the same term in __constant__
// memory and serves every run-time indexed list (a baked object indexed at run time would be a global-memory load).
__device__ __forceinline__ float reward_term(const RlRewardTerm& t, const RlRewardTerm& tc, const Scalars& S, const Layout& L,
                                             const float* sm, const int e, const EnvCtx& c, const int lo, const int hi) {
  const int J = S.num_joints;
  const float* h = sm + L.hist + e * L.hist_pitch;
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
      _Pragma("unroll 1")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) { const float v = SMF(off, j); s += v * v; }
      return s;
    }
    case RL_REW_JOINT_DEVIATION_L1: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jpos, j) - CJ(0, j));
      return s;
    }
    case RL_REW_JOINT_POS_LIMITS: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) {
          const float q = SMF(L.jpos, j);
          float o = -fminf(q - CJ(2, j), 0.f);
          o += fmaxf(q - CJ(3, j), 0.f);
          s += o;
        }
      return s;
    }
    case RL_REW_JOINT_VEL_LIMITS: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += clampf(fabsf(SMF(L.jvel, j)) - CJ(4, j) * t.p[0], 0.f, 1.f);
      return s;
    }
    case RL_REW_JOINT_POWER: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jvel, j) * SMF(L.jtau, j));
      return s;
    }
    case RL_REW_STAND_STILL: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jpos, j) - CJ(0, j));
      s *= (c.cmd_norm < t.p[0]) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_JOINT_POS_PENALTY: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) { const float d = SMF(L.jpos, j) - CJ(0, j); s += d * d; }
      const float running = sqrtf(s);
      const bool moving = (c.cmd_norm > t.p[2]) || (c.vxy_norm > t.p[1]);
      return (moving ? running : t.p[0] * running) * c.gate;
    }
    case RL_REW_JOINT_MIRROR: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const float d = SMF(L.jpos, tc.idx_a[i]) - SMF(L.jpos, tc.idx_b[i]);
        s += d * d;
      }
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_MIRROR: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const float d = fabsf(SMF(L.act, tc.idx_a[i])) - fabsf(SMF(L.act, tc.idx_b[i]));
        s += d * d;
      }
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_SYNC: {
      float r = 0.f;
      _Pragma("unroll 1")
      for (int g = 0; g < t.n_idx; ++g) {
        const int start = tc.idx_b[g], n = tc.idx_c[g];
        if (n < 2) continue;
        float m = 0.f;
        for (int i = 0; i < n; ++i) m += fabsf(SMF(L.act, tc.idx_a[start + i]));
        m = m / (float)n;
        float v = 0.f;
        for (int i = 0; i < n; ++i) { const float d = fabsf(SMF(L.act, tc.idx_a[start + i])) - m; v += d * d; }
        r += v / (float)n;
      }
      return (r * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_RATE_L2: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int a = 0; a < S.n_actions; ++a) { const float d = SMF(L.act, a) - SMF(L.pact, a); s += d * d; }
      return s;
    }
    case RL_REW_UNDESIRED_CONTACTS: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int b = lo; b < S.num_hist_bodies && b < hi; ++b)
        if (((t.body_mask >> b) & 1ull) && (HIST_MAX_NORM(h, b) > t.p[0])) s += 1.f;
      return s * c.gate;   // gate distributes over the two halves of a split term
    }
    case RL_REW_CONTACT_FORCES: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int b = lo; b < S.num_hist_bodies && b < hi; ++b)
        if ((t.body_mask >> b) & 1ull) s += fmaxf(HIST_MAX_NORM(h, b) - t.p[0], 0.f);
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
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = tc.idx_a[i];
        s += (SMF(L.lair, b) - t.p[0]) * (first_contact(sm, L, S, e, b) ? 1.f : 0.f);
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: {
      int n_contact = 0;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) n_contact += (SMF(L.ccon, tc.idx_a[i]) > 0.f) ? 1 : 0;
      const bool single = (n_contact == 1);
      float r = INFINITY;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = tc.idx_a[i];
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
      _Pragma("unroll 1")
      for (int which = 0; which < 2; ++which) {
        const int off = which == 0 ? L.lair : L.lcon;
        float mean = 0.f, m2 = 0.f;
        _Pragma("unroll 1")
        for (int i = 0; i < t.n_idx; ++i) {
          const float x = fminf(SMF(off, tc.idx_a[i]), 0.5f);
          const float d = x - mean;
          mean += d / (float)(i + 1);
          m2 += d * (x - mean);
        }
        r += m2 / (float)(t.n_idx - 1);
      }
      return r * c.gate;
    }
    case RL_REW_FEET_GAIT: {
      const int f00 = tc.idx_a[0], f01 = tc.idx_a[1], f10 = tc.idx_a[2], f11 = tc.idx_a[3];
      const float me2 = t.p[1], sd = t.p[0];
      auto sync = [&](int a, int b) {
        const float da = SMF(L.cair, a) - SMF(L.cair, b);
        const float dc = SMF(L.ccon, a) - SMF(L.ccon, b);
        return rl_expf(-(fminf(da * da, me2) + fminf(dc * dc, me2)) / sd);
      };
      auto async = [&](int a, int b) {
        const float d0 = SMF(L.cair, a) - SMF(L.ccon, b);
        const float d1 = SMF(L.ccon, a) - SMF(L.cair, b);
        return rl_expf(-(fminf(d0 * d0, me2) + fminf(d1 * d1, me2)) / sd);
      };
      const float sync_r = sync(f00, f01) * sync(f10, f11);
      const float async_r = ((async(f00, f10) * async(f01, f11)) * async(f00, f11)) * async(f10, f01);
      const bool moving = (c.cmd_norm > t.p[3]) || (c.vxy_norm > t.p[2]);
      return (moving ? sync_r * async_r : 0.f) * c.gate;
    }
    case RL_REW_FEET_CONTACT: {
      int n = 0;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) n += first_contact(sm, L, S, e, tc.idx_a[i]) ? 1 : 0;
      float r = ((float)n != t.p[0]) ? 1.f : 0.f;
      r *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_CONTACT_WITHOUT_CMD: {
      int n = 0;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) n += first_contact(sm, L, S, e, tc.idx_a[i]) ? 1 : 0;
      float r = (float)n;
      r *= (c.cmd_norm < 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_STUMBLE: {
      bool any = false;   // t = 0 is the newest history sample = net_forces_w
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = tc.idx_c[i];
        const float fx = h[3 * b + 0], fy = h[3 * b + 1], fz = h[3 * b + 2];
        any = any || (sqrtf(fx * fx + fy * fy) > 4.f * fabsf(fz));
      }
      return (any ? 1.f : 0.f) * c.gate;
    }
    case RL_REW_FEET_SLIDE: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int i = lo; i < t.n_idx && i < hi; ++i) {   // [lo, hi): this part's slice of the feet list
        const V3 vw = body_vec(sm, L.bvel, e, tc.idx_b[i]);
        const V3 vb = quat_apply_inverse(c.qw, c.q, V3{vw.x - c.vw.x, vw.y - c.vw.y, vw.z - c.vw.z});
        const float lat = sqrtf(vb.x * vb.x + vb.y * vb.y);
        s += lat * ((HIST_MAX_NORM(h, tc.idx_c[i]) > 1.0f) ? 1.f : 0.f);
      }
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 p = body_vec(sm, L.bpos, e, tc.idx_b[i]);
        const V3 v = body_vec(sm, L.bvel, e, tc.idx_b[i]);
        const float d = p.z - t.p[0];
        s += (d * d) * rl_tanhf(t.p[1] * sqrtf(v.x * v.x + v.y * v.y));
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT_BODY: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec(sm, L.bpos, e, tc.idx_b[i]);
        const V3 vw = body_vec(sm, L.bvel, e, tc.idx_b[i]);
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
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec(sm, L.bpos, e, tc.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const float want = (t.p[0] / 2.f) * ((i % 2 == 0) ? 1.f : -1.f);
        const float d = want - pb.y;
        s += d * d;
      }
      return rl_expf(-s / t.p[1]) * c.gate;
    }
    case RL_REW_FEET_DISTANCE_XY_EXP: {
      float s = 0.f;
      _Pragma("unroll 1")
      for (int i = 0; i < 4; ++i) {
        const V3 pw = body_vec(sm, L.bpos, e, tc.idx_b[i]);
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
      _Pragma("unroll 1")
      for (int i = 0; i < t.n_idx; ++i) {
        const float jv = fabsf(SMF(L.jvel, tc.idx_b[i]));
        const float ta = SMF(L.cair, tc.idx_a[i]);
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

__device__ __forceinline__ EnvCtx make_ctx(const float* sm, const Layout& L, int e) {
  EnvCtx c;
  c.qw = SMF(L.quat, 0);
  c.q = V3{SMF(L.quat, 1), SMF(L.quat, 2), SMF(L.quat, 3)};
  c.pos = V3{SMF(L.root_pos, 0), SMF(L.root_pos, 1), SMF(L.root_pos, 2)};
  c.vw = V3{SMF(L.lin_vel, 0), SMF(L.lin_vel, 1), SMF(L.lin_vel, 2)};
  c.ww = V3{SMF(L.ang_vel, 0), SMF(L.ang_vel, 1), SMF(L.ang_vel, 2)};
#if RL_SHARED_CTX
  c.g = V3{SMF(L.ctxv, 0), SMF(L.ctxv, 1), SMF(L.ctxv, 2)};     // published by warps 0-2 in front of stage 1
  c.vb = V3{SMF(L.ctxv, 3), SMF(L.ctxv, 4), SMF(L.ctxv, 5)};
  c.wb = V3{SMF(L.ctxv, 6), SMF(L.ctxv, 7), SMF(L.ctxv, 8)};
#else
  c.g = quat_apply_inverse(c.qw, c.q, V3{0.f, 0.f, -1.f});
  c.vb = quat_apply_inverse(c.qw, c.q, c.vw);
  c.wb = quat_apply_inverse(c.qw, c.q, c.ww);
#endif
  c.gate = clampf(-c.g.z, 0.f, 0.7f) / 0.7f;
  c.c0 = SMF(L.cmd, 0); c.c1 = SMF(L.cmd, 1); c.c2 = SMF(L.cmd, 2);
  c.cmd_norm = sqrtf((c.c0 * c.c0 + c.c1 * c.c1) + c.c2 * c.c2);
  c.vxy_norm = sqrtf(c.vb.x * c.vb.x + c.vb.y * c.vb.y);
  c.terminated = false;
  return c;
}

// CommandTerm.compute [IL] + UniformThresholdVelocityCommand (V/mdp/commands.py:43-85; the "pits" branch
// is identically off for the in-scope terrains, V/mdp/utils.py:27-28). The new command goes to the cmdn slot
// (other warps still read the old one in this stage); timers, flags and metrics are updated in place.
__device__ __forceinline__ void command_update(float* sm, const Layout& L, const Scalars& S, const RlCommandCfg& cc,
                                               const KArgs& a, const RandState rs, int e, long long env,
                                               const EnvCtx& c, bool write) {
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
      const uint4 r0 = rl_philox(rs, env, RL_STREAM_COMMAND, 0), r1 = rl_philox(rs, env, RL_STREAM_COMMAND, 1);
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
    const float heading = rl_atan2f(fwd.y, fwd.x);
    const float err = wrap_to_pi(head - heading);
    c2 = clampf(cc.heading_control_stiffness * err, cc.ang_vel_z_lo, cc.ang_vel_z_hi);
  }
  if (isstand) { c0 = 0.f; c1 = 0.f; c2 = 0.f; }
  if (!write) { c0 = c.c0; c1 = c.c1; c2 = c.c2; }
  SMF(L.cmdn, 0) = c0; SMF(L.cmdn, 1) = c1; SMF(L.cmdn, 2) = c2;
  if (write) {
    SMF(L.tleft, 0) = tleft; SMF(L.head, 0) = head;
    SMF(L.ishead, 0) = __int_as_float(ishead); SMF(L.isstand, 0) = __int_as_float(isstand);
  }
}

// Columns [lo, hi) of one observation term for env e: ObservationManager.compute_group [IL]
// (clone -> +noise -> clip -> scale), written into the group's shared-memory row.
__device__ __forceinline__ void obs_task(float* sm, const Layout& L, const Scalars& S, const RlObsTerm& t, const RlObsTerm& tc,
                                         const bool corrupt, const KArgs& a, const RandState rs, const int g,
                                         const int ti, const int col0, const int lo, const int hi, const int e,
                                         const long long env, const EnvCtx& c, const int eplen_now) {
  float* row = sm + LOBS(g) + e * LOBSP(g);
  // noise-as-input mode (RlRandom.obs_uniforms: reproducibility hook for tests / replays, not the production
  // path): read straight from global memory
  const float* urow = a.rnd.obs_uniforms[g] ? a.rnd.obs_uniforms[g] + env * (g == 0 ? S.obs_dim0 : S.obs_dim1) : nullptr;
  const bool ext_u = (a.rnd.obs_uniforms[g] != nullptr);
  const bool noisy = t.has_noise && corrupt;
  _Pragma("unroll 1")
  for (int qd = lo / 4; qd * 4 < hi; ++qd) {
    float u4[4] = {0.f, 0.f, 0.f, 0.f};
    if (noisy && !ext_u) {
      const uint4 r = rl_philox(rs, env, RL_STREAM_OBS + g * RL_MAX_OBS_TERMS + ti, (uint32_t)qd);
      u4[0] = u01(r.x); u4[1] = u01(r.y); u4[2] = u01(r.z); u4[3] = u01(r.w);
    }
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
      const int col = qd * 4 + r4;
      if (col >= hi) break;
      float v;
      switch (t.type) {
        case RL_OBS_BASE_LIN_VEL: v = col == 0 ? c.vb.x : (col == 1 ? c.vb.y : c.vb.z); break;
        case RL_OBS_BASE_ANG_VEL: v = col == 0 ? c.wb.x : (col == 1 ? c.wb.y : c.wb.z); break;
        case RL_OBS_PROJECTED_GRAVITY: v = col == 0 ? c.g.x : (col == 1 ? c.g.y : c.g.z); break;
        case RL_OBS_GENERATED_COMMANDS: v = SMF(L.cmdn, col); break;
        case RL_OBS_JOINT_POS_REL: v = SMF(L.jpos, tc.ids[col]) - CJ(0, tc.ids[col]); break;
        case RL_OBS_JOINT_POS_REL_WITHOUT_WHEEL:
          v = SMF(L.jpos, tc.ids[col]) - CJ(0, tc.ids[col]);
          if ((t.zero_mask >> col) & 1ull) v = 0.f;
          break;
        case RL_OBS_JOINT_VEL_REL: v = SMF(L.jvel, tc.ids[col]) - CJ(1, tc.ids[col]); break;
        case RL_OBS_LAST_ACTION: v = SMF(L.act, col); break;
        case RL_OBS_HEIGHT_SCAN: v = (SMF(L.raypos, 0) - sm[L.rays + e * L.rays_pitch + col]) - t.p[0]; break;
        case RL_OBS_PHASE: {
          const float ph = ((float)eplen_now * S.step_dt) / t.p[0];
          v = col == 0 ? rl_sinf((2.f * RL_PI_F) * ph) : rl_cosf((2.f * RL_PI_F) * ph);
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
}


// ---------------------------------------------------------------------------------------------------
// Tile movers (one code copy each)
// ---------------------------------------------------------------------------------------------------
// The small per-env fields: field f is moved by warp f mod NW, its lanes stride over the field's rows (a "row" is
// one component of one field; record word w of local env e lives at sm[w*kE + e]). Deliberately a LOOP over
// per-field metadata and not per-field code: every SM sees this code exactly once per launch, so its size - not
// its instruction count - is what the load phase costs (profiles/r1_front_end.md).
__device__ __noinline__ void field_load_elems(float* sm, FieldD fd, int nc, int w0, int env0, int nvalid,
                                              const int32_t* ids, int lane) {
  if (lane >= nvalid) return;   // any strides, ragged tiles, env-id lists: lane = env, one element per copy
  const long long env = ids ? (long long)ids[env0 + lane] : (long long)(env0 + lane);
  const float* src = static_cast<const float*>(fd.ptr) + env * fd.es;
  float* dst = sm + w0 * kE + lane;
#pragma unroll 4
  for (int c = 0; c < nc; ++c) cp_async4(dst + c * kE, src + (long long)c * fd.cs);
}
__device__ __noinline__ void field_store_elems(const float* sm, FieldD fd, int nc, int w0, int env0, int nvalid,
                                               const int32_t* ids, int lane) {
  if (lane >= nvalid) return;
  const long long env = ids ? (long long)ids[env0 + lane] : (long long)(env0 + lane);
  float* dst = static_cast<float*>(const_cast<void*>(fd.ptr)) + env * fd.es;
  const float* src = sm + w0 * kE + lane;
#pragma unroll 4
  for (int c = 0; c < nc; ++c) dst[(long long)c * fd.cs] = src[c * kE];
}
__device__ __forceinline__ void load_fields(float* sm, const KArgs& a, int env0, int nvalid, const int32_t* ids,
                                            bool full, int warp, int lane, int nwarps) {
#pragma unroll 1
  for (int f = warp; f < IF_COUNT; f += nwarps) {
    if (!((a.in_mask >> f) & 1u)) continue;
    const FieldD fd = a.in[f];
    const int nc = fd.meta & 0xffff, w0 = fd.meta >> 16;
    if (full && ((a.in_vec4 >> f) & 1u)) {   // SoA rows: 16 bytes (4 envs) per copy
      const float* src = static_cast<const float*>(fd.ptr) + env0;
#pragma unroll 1
      for (int j = lane; j < nc * (kE / 4); j += 32)
        cp_async16(sm + (w0 + (j >> 3)) * kE + 4 * (j & 7), src + (size_t)(j >> 3) * fd.cs + 4 * (j & 7));
    } else {
      field_load_elems(sm, fd, nc, w0, env0, nvalid, ids, lane);
    }
  }
}
__device__ __forceinline__ void store_fields(const float* sm, const KArgs& a, uint32_t omask, int env0, int nvalid,
                                             const int32_t* ids, bool full, int warp, int lane, int nwarps) {
#pragma unroll 1
  for (int f = warp; f < OF_COUNT; f += nwarps) {
    if (!((omask >> f) & 1u)) continue;
    const FieldD fd = a.outf[f];
    const int nc = fd.meta & 0xffff, w0 = fd.meta >> 16;
    if (full && ((a.out_vec4 >> f) & 1u)) {
      float* dst = static_cast<float*>(const_cast<void*>(fd.ptr)) + env0;
#pragma unroll 1
      for (int j = lane; j < nc * (kE / 4); j += 32)
        *reinterpret_cast<float4*>(dst + (size_t)(j >> 3) * fd.cs + 4 * (j & 7)) =
            *reinterpret_cast<const float4*>(sm + (w0 + (j >> 3)) * kE + 4 * (j & 7));
    } else {
      field_store_elems(sm, fd, nc, w0, env0, nvalid, ids, lane);
    }
  }
}

// AoS span [kE][pitch] <-> global rows, element-wise (used when a span is not one aligned contiguous block or the
// shared-memory pitch is padded)
__device__ __noinline__ void span_load_elems(float* dst, int pitch, FieldD fd, int ncomp, int env0, int nvalid,
                                             const int32_t* ids, int tid, int nthreads) {
  const int warp = tid >> 5, lane = tid & 31, nw = nthreads >> 5;
  for (int el = warp; el < nvalid; el += nw) {   // a warp walks one env's row: coalesced when rows are contiguous
    const long long env = ids ? (long long)ids[env0 + el] : (long long)(env0 + el);
    const float* src = static_cast<const float*>(fd.ptr) + env * fd.es;
    for (int c = lane; c < ncomp; c += 32) cp_async4(dst + el * pitch + c, src + (long long)c * fd.cs);
  }
}
__device__ __noinline__ void span_store_elems(const float* src, int pitch, float* ptr, long long es, int ncomp,
                                              int env0, int nvalid, const int32_t* ids, int tid, int nthreads) {
  const int warp = tid >> 5, lane = tid & 31, nw = nthreads >> 5;
  for (int el = warp; el < nvalid; el += nw) {
    const long long env = ids ? (long long)ids[env0 + el] : (long long)(env0 + el);
    for (int c = lane; c < ncomp; c += 32) ptr[env * es + c] = src[el * pitch + c];
  }
}
// one aligned contiguous [kE][ncomp] block in global memory AND an unpadded shared-memory pitch
__device__ __forceinline__ bool span_bulk_ok(const FieldD& fd, int ncomp, int pitch, int env0) {
  if (fd.ptr == nullptr || ncomp <= 0 || pitch != ncomp || fd.cs != 1 || fd.es != ncomp) return false;
  const uintptr_t p = reinterpret_cast<uintptr_t>(fd.ptr) + (uintptr_t)env0 * (uintptr_t)ncomp * 4u;
  return ((p & 15u) == 0) && ((((long long)kE * ncomp * 4) & 15) == 0);
}
__device__ __forceinline__ void store_u8(const float* sm, int off, const FieldD& fd, int env0, int nvalid,
                                         const int32_t* ids, int tid) {
  if (fd.ptr != nullptr && tid < nvalid) {
    const long long env = ids ? (long long)ids[env0 + tid] : (long long)(env0 + tid);
    static_cast<uint8_t*>(const_cast<void*>(fd.ptr))[env * fd.es] = (uint8_t)__float_as_int(sm[off + tid]);
  }
}

// ---------------------------------------------------------------------------------------------------
// The fused step kernel: kE = 32 envs per CTA, NW warps. MODE 0 = step, MODE 1 = single-term evaluation.
//   load   : everything asynchronous (TMA bulk copies + cp.async), one join
//   stage 1: every warp runs its share of the schedule, thread-per-env (lane e = env e)
//   stage 2: warp 0 adds the reward up in manager order (and finishes is_terminated)
//   store  : bulk stores for the observation rows, per-field loops for the SoA outputs; last CTA compacts reset ids
// ---------------------------------------------------------------------------------------------------
// the variants keep the second resident CTA per SM (<= 64 registers at 512 threads): without the hint ptxas took 104
#if RL_SHARED_NORMS || RL_SHARED_CTX || RL_PERSISTENT
#define RL_STEP_BOUNDS __launch_bounds__(NW * 32 * TILES, (P::kStatic && NW * 32 * TILES <= 512) ? 2 : 1)   /* the generic kernel keeps its registers */
#else
#define RL_STEP_BOUNDS __launch_bounds__(NW * 32 * TILES)
#endif
template <class P, int NW, int MODE, bool DBG, int TILES>
__global__ void RL_STEP_BOUNDS mdp_step_kernel(const KArgs a) {
  extern __shared__ __align__(128) float sm_cta[];
  __shared__ __align__(8) uint64_t s_bar_all[TILES];
  __shared__ int s_last_all[TILES];
  const Scalars S = P::scalars(a);
  const Layout L = P::layout(a);
  constexpr int NT = NW * 32;   // threads per tile
  // TILES > 1: the CTA carries TILES independent tiles; threads [t*NT, (t+1)*NT) are "virtual CTA" vb of tile t with
  // their own record, mbarrier and named barrier, so warp w of every tile runs the same task code at about the same
  // time (one instruction fetch per SM serves all of them). Everything below is written in terms of the virtual CTA.
  const int tile = (TILES > 1) ? (int)(threadIdx.x / NT) : 0;
  const int tid = (TILES > 1) ? (int)(threadIdx.x - tile * NT) : (int)threadIdx.x;
#if RL_PERSISTENT
  const int vb_first = (int)(blockIdx.x * TILES + tile);
  const int vgrid = a.vgrid;
#else
  const int vb = (TILES > 1) ? (int)(blockIdx.x * TILES + tile) : (int)blockIdx.x;        // virtual CTA = tile index
  const int vgrid = (TILES > 1) ? a.vgrid : (int)gridDim.x;
  if (TILES > 1 && vb >= vgrid) return;   // odd tile count: the spare half of the last CTA
#endif
  float* const sm = sm_cta + (TILES > 1 ? (size_t)tile * (size_t)a.tile_words : (size_t)0);
  uint64_t& s_bar = s_bar_all[tile];
  int& s_last = s_last_all[tile];
  auto tile_sync = [&]() __attribute__((always_inline)) {
    if constexpr (TILES > 1) asm volatile("bar.sync %0, %1;" ::"r"(tile + 1), "n"(NT) : "memory");
    else __syncthreads();
  };
  auto tile_sync_or = [&](bool pred) __attribute__((always_inline)) -> int {
    if constexpr (TILES > 1) {
      uint32_t r;
      asm volatile(
          "{\n"
          ".reg .pred p, q;\n"
          "setp.ne.u32 q, %3, 0;\n"
          "bar.red.or.pred p, %1, %2, q;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(r)
          : "r"(tile + 1), "n"(NT), "r"((uint32_t)pred)
          : "memory");
      return (int)r;
    } else {
      return __syncthreads_or(pred);
    }
  };
#if RL_PERSISTENT
  for (int vb_it = vb_first; vb_it < vgrid; vb_it += (int)gridDim.x * TILES) {
  const int vb = vb_it;   // the tile of this iteration; the body below is the one-tile kernel, "return" = next tile
  [&]() __attribute__((always_inline)) {
#endif
#define RL_STAMP(i) do { if constexpr (DBG) { if (tid == 0) a.dbg[(size_t)vb * RL_DEBUG_STRIDE + (i)] = clock64(); } } while (0)
#define RL_SUB(i) RL_STAMP(8 + RL_MAX_TASKS + 32 + (i))   /* finer stamps inside the load phase (debug build only) */
  RL_STAMP(0);
  const int warp = tid >> 5;
  const int e = tid & 31;       // compute phase: this lane's env inside the tile
  const uint32_t ph = a.phases;
  if (a.use_pdl) {
    // launch-latency overlap only: every read below may depend on the predecessor, so wait first
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int n_total = a.has_ids ? *a.n_env_ids : a.N;
  const int env0 = vb * kE;
  const int K = S.num_reward_terms;
  // RESET comes in two forms: on an env-id list (gathered tiles, every env of the launch is reset) or - without
  // a list - over all envs, resetting those whose terminated | truncated byte is set (full-tile fast path)
  const bool do_reset = (MODE == 0) && (ph & RL_PHASE_RESET) != 0;
  const bool reset_masked = do_reset && !a.has_ids;
  // (the masked form reads its reset count only in the tail: a dependent global load up here would stall the
  // whole prologue; the id-list form needs its count for the bounds anyway)
  if (do_reset && !reset_masked && n_total == 0 && vb == 0 && tid < RL_LOG_STRIDE) {
    // nothing to reset and no CTA reaches the tail: the logged scalars are defined as 0
    if (tid < K) { if (a.out.reset_log.episode_sum_mean) a.out.reset_log.episode_sum_mean[tid] = 0.f; }
    else if (tid < K + RL_MAX_DONE_TERMS) { if (a.out.reset_log.done_term_count) a.out.reset_log.done_term_count[tid - K] = 0.f; }
    else if (tid < K + RL_MAX_DONE_TERMS + 2) { if (a.out.reset_log.metric_mean) a.out.reset_log.metric_mean[tid - K - RL_MAX_DONE_TERMS] = 0.f; }
  }
  if (env0 >= n_total && !(ph & RL_PHASE_COMPACT)) return;
  const int nvalid = max(0, min(kE, n_total - env0));
  const int32_t* ids = a.has_ids ? a.env_ids : nullptr;
  const int J = S.num_joints, A = S.n_actions;
  const int R = S.num_rays;
  const int HW = S.hist_len * S.num_hist_bodies * 3;
  const bool need_hist = (MODE == 1) || (ph & (RL_PHASE_DONES | RL_PHASE_REWARDS));
  const bool need_rays = (MODE == 0) && (ph & RL_PHASE_OBS) && R > 0 && a.rays.ptr != nullptr;
  const bool full = (nvalid == kE) && (ids == nullptr);
  RandState rs;
  rs.seed = a.rnd.seed;
  rs.step = a.rnd.step + (a.rnd.step_counter ? *a.rnd.step_counter : 0ull);
  rs.env_id_offset = a.rnd.env_id_offset;

  // ---- load phase: everything is asynchronous, nothing below waits until the single join point ---------
  if (nvalid > 0) {
    // the big transfers first: their DRAM latency overlaps everything else the prologue does
    const bool hist_bulk = need_hist && full && span_bulk_ok(a.hist, HW, L.hist_pitch, env0);
    const bool rays_bulk = need_rays && full && span_bulk_ok(a.rays, R, L.rays_pitch, env0);
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      uint32_t bytes = 0;
      if (hist_bulk) bytes += (uint32_t)(kE * HW * 4);
      if (rays_bulk) bytes += (uint32_t)(kE * R * 4);
      mbar_expect_tx(&s_bar, bytes);
      if (hist_bulk) bulk_g2s(sm + L.hist, static_cast<const float*>(a.hist.ptr) + (size_t)env0 * HW, (uint32_t)(kE * HW * 4), &s_bar);
      if (rays_bulk) bulk_g2s(sm + L.rays, static_cast<const float*>(a.rays.ptr) + (size_t)env0 * R, (uint32_t)(kE * R * 4), &s_bar);
    }
    // byte-sized per-env flags: plain loads into registers, issued early so that their latency overlaps the
    // issue of everything else
    const bool want_cmd_flags = (MODE == 0) && (ph & (RL_PHASE_COMMAND | RL_PHASE_RESET)) != 0;
    const bool want_done_bits = do_reset && a.out.done_bits != nullptr;
    int u8_head = 0, u8_stand = 0, u8_bits = 0, u8_reset = do_reset ? 1 : 0;
    if (tid < nvalid) {
      const long long ev = ids ? (long long)ids[env0 + tid] : (long long)(env0 + tid);
      if (reset_masked) u8_reset = (a.out.terminated[ev] | a.out.truncated[ev]) != 0;
      if (want_cmd_flags) {
        if (a.is_heading.ptr) u8_head = static_cast<const uint8_t*>(a.is_heading.ptr)[ev * a.is_heading.es];
        if (a.is_standing.ptr) u8_stand = static_cast<const uint8_t*>(a.is_standing.ptr)[ev * a.is_standing.es];
      }
      if (want_done_bits) u8_bits = a.out.done_bits[ev];
    }
    RL_SUB(0);                  // prologue + bulk copies issued
    load_fields(sm, a, env0, nvalid, ids, full, warp, e, NW);
    RL_SUB(1);                  // per-field copies issued
    if (need_hist && !hist_bulk && a.hist.ptr) span_load_elems(sm + L.hist, L.hist_pitch, a.hist, HW, env0, nvalid, ids, tid, NT);
    if (need_rays && !rays_bulk) span_load_elems(sm + L.rays, L.rays_pitch, a.rays, R, env0, nvalid, ids, tid, NT);
    if (ph & RL_PHASE_REWARDS)
      for (int i = tid; i < K * kE; i += NT) {
        sm[L.arrive + i] = __int_as_float(0);
        if ((a.rw_zero >> (i / kE)) & 1ull) {   // weight-0 terms: no task evaluates them
          sm[L.stepr + i] = 0.f;
          sm[L.termv + (i / kE) * kTermParts * kE + (i % kE)] = 0.f;
        }
      }
    RL_SUB(2);                  // span loads issued
    // per-joint constants: constant bank -> shared
    for (int i = tid; i < J; i += NT)
      P::for_joint_consts(a, i, [&](float q0, float qd0, float lo, float hi, float vl) {
        sm[L.cj + 0 * J + i] = q0; sm[L.cj + 1 * J + i] = qd0; sm[L.cj + 2 * J + i] = lo;
        sm[L.cj + 3 * J + i] = hi; sm[L.cj + 4 * J + i] = vl;
      });
    if (tid < nvalid) {         // the byte flags were requested at the top of the load phase
      if (want_cmd_flags) { sm[L.ishead + tid] = __int_as_float(u8_head); sm[L.isstand + tid] = __int_as_float(u8_stand); }
      if (want_done_bits) sm[L.flags + tid] = __int_as_float(u8_bits);
    }
    if (do_reset && tid < kE) sm[L.rmask + tid] = __int_as_float(tid < nvalid ? u8_reset : 0);
    RL_STAMP(1);                // all loads issued
    cp_async_wait_all();
    tile_sync();            // record + mbarrier init visible to everyone
    mbar_wait(&s_bar, 0);       // bulk copies landed
    RL_STAMP(2);                // tile resident
  }

  // Tickets of the two "last CTA" tails are taken EARLY - as soon as what the last CTA will read has been written -
  // and only looked at in the tail: the fence + atomic round trip (~1 us) overlaps the rest of the tile's work.
  unsigned early_prev = 0;
  const bool compact = (ph & RL_PHASE_COMPACT) && (ph & RL_PHASE_DONES) && !a.has_ids;
  constexpr int kCompactWarp = 1;   // idle during stage 2 (warp 0 sums the reward up)
  auto compact_arrive = [&]() __attribute__((always_inline)) {   // one warp: the tile's done mask, then its ticket
    const int f2 = (e < nvalid) ? __float_as_int(sm[L.flags + e]) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, (f2 >> 8) & 3);
    if (e == 0) {
      a.cta_mask[vb] = m;
      early_prev = ticket_arrive_release(a.ticket);
    }
  };
  const bool valid = e < nvalid;
  const long long env = valid ? (ids ? (long long)ids[env0 + e] : (long long)(env0 + e)) : 0;
  if (nvalid > 0) {
    // ---- manager reset of the tile's envs that are being reset ------------------------------------------
    const int tile_resets = do_reset ? tile_sync_or(tid < kE && __float_as_int(sm[L.rmask + (tid & 31)]) != 0) : 0;
    if (do_reset && !tile_resets) {   // CTA-uniform: nothing to reset here, the logging partials are zero
      if (warp == 0) {
        for (int q = e; q < K + RL_MAX_DONE_TERMS + 2; q += 32) a.log_partials[(size_t)vb * RL_LOG_STRIDE + q] = 0.f;
        __syncwarp();
        // the tile's arrival at the logging reduction, long before the tail needs the answer (see the tail)
        if (e == 0) early_prev = ticket_arrive_release(a.ticket);
      }
    }
    if (tile_resets) {
      const bool rme = __float_as_int(sm[L.rmask + e]) != 0;
      // logging partials of this CTA (combined by the last CTA in a fixed order -> deterministic):
      // quantity q is reduced by warp q mod NW over its lanes (= envs) with a fixed shuffle tree
#pragma unroll 1
      for (int q = warp; q < K + RL_MAX_DONE_TERMS + 2; q += NW) {
        float x = 0.f;
        if (rme) {
          if (q < K) x = sm[L.sums + q * kE + e];
          else if (q < K + RL_MAX_DONE_TERMS)
            x = (a.out.done_bits != nullptr) ? (float)((__float_as_int(sm[L.flags + e]) >> (q - K)) & 1) : 0.f;
          else x = sm[(q == K + RL_MAX_DONE_TERMS ? L.mxy : L.myaw) + e];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (e == 0) a.log_partials[(size_t)vb * RL_LOG_STRIDE + q] = x;
      }
      tile_sync();
      if (tid == 0) early_prev = ticket_arrive_release(a.ticket);   // partials of every warp ordered by the barrier
      // RewardManager / ActionManager / CommandTerm .reset [IL], episode_length_buf = 0
      for (int i = tid; i < kE * K; i += NT) if (__float_as_int(sm[L.rmask + (i & 31)]) != 0) sm[L.sums + i] = 0.f;
      for (int i = tid; i < kE * A; i += NT)
        if (__float_as_int(sm[L.rmask + (i & 31)]) != 0) { sm[L.act + i] = 0.f; sm[L.pact + i] = 0.f; }
      if (tid < kE && __float_as_int(sm[L.rmask + tid]) != 0) {
        const int el = tid;
        sm[L.mxy + el] = 0.f; sm[L.myaw + el] = 0.f;
        sm[L.eplen + el] = __int_as_float(0); sm[L.epnew + el] = __int_as_float(0);
        const auto& cc = P::command(a);
        const long long ev = ids ? (long long)ids[env0 + min(el, nvalid - 1)] : (long long)(env0 + el);
        float u[RL_NUM_CMD_UNIFORMS];
        if (a.rnd.cmd_uniforms != nullptr) {
#pragma unroll
          for (int q = 0; q < RL_NUM_CMD_UNIFORMS; ++q) u[q] = sm[L.cmdu + q * kE + el];
        } else {
          const uint4 r0 = rl_philox(rs, ev, RL_STREAM_RESET_COMMAND, 0), r1 = rl_philox(rs, ev, RL_STREAM_RESET_COMMAND, 1);
          u[0] = u01(r0.x); u[1] = u01(r0.y); u[2] = u01(r0.z); u[3] = u01(r0.w);
          u[4] = u01(r1.x); u[5] = u01(r1.y); u[6] = u01(r1.z);
        }
        float c0 = u[1] * (cc.lin_vel_x_hi - cc.lin_vel_x_lo) + cc.lin_vel_x_lo;
        float c1 = u[2] * (cc.lin_vel_y_hi - cc.lin_vel_y_lo) + cc.lin_vel_y_lo;
        const float c2 = u[3] * (cc.ang_vel_z_hi - cc.ang_vel_z_lo) + cc.ang_vel_z_lo;
        const float keep = (sqrtf(c0 * c0 + c1 * c1) > cc.small_cmd_threshold) ? 1.f : 0.f;
        c0 *= keep; c1 *= keep;
        sm[L.cmd + 0 * kE + el] = c0; sm[L.cmd + 1 * kE + el] = c1; sm[L.cmd + 2 * kE + el] = c2;
        sm[L.cmdn + 0 * kE + el] = c0; sm[L.cmdn + 1 * kE + el] = c1; sm[L.cmdn + 2 * kE + el] = c2;
        sm[L.tleft + el] = u[0] * (cc.resampling_time_hi - cc.resampling_time_lo) + cc.resampling_time_lo;
        if (cc.heading_command) {
          sm[L.head + el] = u[4] * (cc.heading_hi - cc.heading_lo) + cc.heading_lo;
          sm[L.ishead + el] = __int_as_float((u[5] <= cc.rel_heading_envs) ? 1 : 0);
        }
        sm[L.isstand + el] = __int_as_float((u[6] <= cc.rel_standing_envs) ? 1 : 0);
      }
      tile_sync();
    }

#if RL_SHARED_NORMS || RL_SHARED_CTX
    {  // prepass in front of stage 1: work every warp would otherwise repeat, done once; lane = env
#if RL_SHARED_CTX
      if (warp < 3) {   // NW >= 4
        const float qw = SMF(L.quat, 0);
        const V3 q{SMF(L.quat, 1), SMF(L.quat, 2), SMF(L.quat, 3)};
        const int src = (warp == 1) ? L.lin_vel : L.ang_vel;
        const V3 v = (warp == 0) ? V3{0.f, 0.f, -1.f} : V3{SMF(src, 0), SMF(src, 1), SMF(src, 2)};
        const V3 r = quat_apply_inverse(qw, q, v);
        SMF(L.ctxv, 3 * warp + 0) = r.x; SMF(L.ctxv, 3 * warp + 1) = r.y; SMF(L.ctxv, 3 * warp + 2) = r.z;
      }
#endif
#if RL_SHARED_NORMS
      if (need_hist) {   // contact norms: one code copy for all warps; the warps without a rotation go first
        const float* hrow = sm + L.hist + e * L.hist_pitch;
        const int first = RL_SHARED_CTX ? (warp + NW - 3) % NW : warp;
#pragma unroll 1
        for (int b = first; b < S.num_hist_bodies; b += NW) SMF(L.hnorm, b) = hist_max_norm(hrow, S.hist_len, S.num_hist_bodies, b);
      }
#endif
      tile_sync();
    }
#endif
    // ---- stage 1: thread-per-env, warps run different tasks ------------------------------------------------
    EnvCtx c = make_ctx(sm, L, e);
    if (MODE == 1) {
      if (warp == 0) {
        if (a.ext_terminated != nullptr && valid) c.terminated = a.ext_terminated[env] != 0;
        const float v = reward_term(*a.adhoc, *a.adhoc, S, L, sm, e, c, 0, 64);
        if (valid) a.term_out[env] = v;
      }
      return;
    }
    const int eplen_now = __float_as_int(SMF(L.eplen, 0)) + ((ph & RL_PHASE_DONES) ? 1 : 0);
    // TerminationManager.compute [IL] for this lane's env: bits | terminated << 8 | time_out << 9
    auto eval_dones = [&]() -> int {
      uint32_t bits = 0, term = 0, trunc = 0;
      const float* h = sm + L.hist + e * L.hist_pitch;
      P::for_dones(a, [&](const RlDoneTerm& t, int d) __attribute__((always_inline)) {
        int fired = 0;
        if (t.type == RL_DONE_TIME_OUT) {
          fired = eplen_now >= S.max_episode_length;
        } else if (t.type == RL_DONE_TERRAIN_OUT_OF_BOUNDS) {
          fired = (t.p[2] != 0.f) && ((fabsf(c.pos.x) > t.p[0]) || (fabsf(c.pos.y) > t.p[1]));
        } else if (t.type == RL_DONE_ILLEGAL_CONTACT) {
          _Pragma("unroll 1")
          for (int b = 0; b < S.num_hist_bodies; ++b)
            if (((t.body_mask >> b) & 1ull) && (HIST_MAX_NORM(h, b) > t.p[0])) fired = 1;
        }
        if (fired) { bits |= 1u << d; if (t.time_out) trunc = 1; else term = 1; }
      });
      return (int)(bits | (term << 8) | (trunc << 9));
    };
    if constexpr (DBG) { if (e == 0) a.dbg[(size_t)vb * RL_DEBUG_STRIDE + 8 + RL_MAX_TASKS + warp] = clock64(); }
    P::template for_tasks<NW>(a, warp, [&](auto, const Task& tk, const RlRewardTerm& rt, const RlObsTerm& ot, const bool corrupt,
                                          const int task_idx) __attribute__((always_inline)) {
      long long t_begin = 0;
      if constexpr (DBG) t_begin = clock64();
      [&]() __attribute__((always_inline)) {
      if (tk.kind == TK_REWARD) {
        if (!(ph & RL_PHASE_REWARDS)) return;
        const float raw = reward_term(rt, c_spec[a.slot].rewards[tk.a], S, L, sm, e, c, tk.lo, tk.hi);
        const int k = tk.a;
        float full_raw = raw;
        bool finish = true;
        if (tk.pad) {
          // one part of a split term: publish the partial sum; whoever arrives last (per env) adds the parts in
          // part order and finishes the term - no serial tail for it in stage 2
          const int parts = tk.col0;
          SMF(L.termv, kTermParts * k + tk.b) = raw;
          __threadfence_block();
          const int old = atomicAdd(reinterpret_cast<int*>(sm) + L.arrive + k * kE + e, 1);
          finish = (old == parts - 1);
          if (finish) {
            __threadfence_block();
            full_raw = *(volatile float*)&SMF(L.termv, kTermParts * k);
#pragma unroll
            for (int p = 1; p < kTermParts; ++p)
              if (p < parts) full_raw += *(volatile float*)&SMF(L.termv, kTermParts * k + p);
          }
        }
        if (finish) {
          // RewardManager.compute [IL]: value = func * weight * dt; sums += value; step_reward = value / dt
          const float val = (full_raw * rt.weight) * S.step_dt;
          SMF(L.termv, kTermParts * k) = val;
          SMF(L.sums, k) = SMF(L.sums, k) + val;
          SMF(L.stepr, k) = rl_div(val, S.step_dt);
        }
      } else if (tk.kind == TK_OBS) {
        if (!(ph & RL_PHASE_OBS) || a.out.obs[tk.a] == nullptr) return;
        obs_task(sm, L, S, ot, c_spec[a.slot].obs[tk.a].terms[tk.b], corrupt, a, rs, tk.a, tk.b, tk.col0, tk.lo, tk.hi, e, env, c, eplen_now);
      } else if (tk.kind == TK_DONES) {
        if (!(ph & RL_PHASE_DONES)) return;
        SMF(L.epnew, 0) = __int_as_float(eplen_now);
        SMF(L.flags, 0) = __int_as_float(eval_dones());
      } else {  // TK_COMMAND: CommandManager.compute + the observation columns that show the new command
        if (!(ph & (RL_PHASE_COMMAND | RL_PHASE_OBS))) return;
        if (ph & RL_PHASE_COMMAND) {
          bool skip = false;
          if ((ph & RL_PHASE_SKIP_DONE_ENVS) && (ph & RL_PHASE_DONES)) skip = ((eval_dones() >> 8) & 3) != 0;
          command_update(sm, L, S, P::command(a), a, rs, e, env, c, !skip);
        } else if (!(ph & RL_PHASE_RESET)) {
          SMF(L.cmdn, 0) = c.c0; SMF(L.cmdn, 1) = c.c1; SMF(L.cmdn, 2) = c.c2;
        }
        if (ph & RL_PHASE_OBS) {
          P::for_cmd_obs(a, [&](const RlObsTerm& t, int g, int ti, int col0, bool corr) __attribute__((always_inline)) {
            if (a.out.obs[g] == nullptr) return;
            obs_task(sm, L, S, t, c_spec[a.slot].obs[g].terms[ti], corr, a, rs, g, ti, col0, 0, t.dim, e, env, c, eplen_now);
          });
        }
      }
      }();
      if constexpr (DBG) { if (e == 0) a.dbg[(size_t)vb * RL_DEBUG_STRIDE + 8 + task_idx] = clock64() - t_begin; }
    });
    tile_sync();
    RL_STAMP(3);                // stage 1 done
    if (compact && warp == kCompactWarp) compact_arrive();   // the termination flags are final

    // ---- stage 2: warp 0 finishes the late terms and adds the reward up in manager order ------------------
    if ((ph & RL_PHASE_REWARDS) && warp == 0) {
      // Kept as small as possible: warp 0 runs this alone, right after a barrier, on code no warp has touched -
      // what it costs is its length in instruction-cache lines. Weight-0 terms were zeroed in the load phase.
      float total = 0.f;
      if (a.rw_late != 0) {   // only is_terminated, which needs the termination result, is finished here
        const int fl = (ph & RL_PHASE_DONES) ? __float_as_int(SMF(L.flags, 0)) : 0;
        const bool terminated = ((fl >> 8) & 1) != 0;
#pragma unroll 1
        for (uint64_t m = a.rw_late & ~a.rw_zero; m != 0; m &= m - 1) {
          const int k = __ffsll((long long)m) - 1;
          const float raw = ((a.rw_isterm >> k) & 1ull) ? (terminated ? 1.f : 0.f) : SMF(L.termv, kTermParts * k);
          const float val = (raw * a.rw_weight[k]) * S.step_dt;
          SMF(L.termv, kTermParts * k) = val;
          SMF(L.sums, k) = SMF(L.sums, k) + val;
          SMF(L.stepr, k) = rl_div(val, S.step_dt);
        }
      }
      RL_SUB(3);                // late terms finished
#pragma unroll 1
      for (int k = 0; k < K; ++k) total += SMF(L.termv, kTermParts * k);   // manager order
      SMF(L.rew, 0) = total;
      RL_SUB(4);                // reward summed
    }
    tile_sync();
    RL_STAMP(4);                // stage 2 done

    // ---- store phase ----------------------------------------------------------------------------
    if (ph & RL_PHASE_OBS) {
      fence_proxy_async();
      tile_sync();
#pragma unroll
      for (int g = 0; g < RL_NUM_OBS_GROUPS; ++g) {
        const int D = P::obs_dim(a, g);
        if (a.out.obs[g] == nullptr || D <= 0) continue;
        const FieldD od{a.out.obs[g], (int)a.out.obs_pitch[g], 1};
        if (full && span_bulk_ok(od, D, LOBSP(g), env0)) {
          if (tid == 0) bulk_s2g(a.out.obs[g] + (size_t)env0 * D, sm + LOBS(g), (uint32_t)(kE * D * 4));
        } else {
          span_store_elems(sm + LOBS(g), LOBSP(g), a.out.obs[g], a.out.obs_pitch[g], D, env0, nvalid, ids, tid, NT);
        }
      }
      if (tid == 0) bulk_commit();
    }
    {
      uint32_t omask = a.out_mask;
      if (do_reset && !tile_resets) {   // nothing was reset here: those fields are unchanged
        omask &= ~((1u << OF_SUMS) | (1u << OF_ACT) | (1u << OF_PACT));
        if (!(ph & RL_PHASE_DONES)) omask &= ~(1u << OF_EPLEN);
      }
      store_fields(sm, a, omask, env0, nvalid, ids, full, warp, e, NW);
    }
    if (ph & RL_PHASE_DONES) {
      if (tid < nvalid) {
        const int f2 = __float_as_int(sm[L.flags + tid]);
        const long long ev = ids ? (long long)ids[env0 + tid] : (long long)(env0 + tid);
        if (a.out.done_bits) a.out.done_bits[ev] = (uint8_t)(f2 & 0xff);
        if (a.out.terminated) a.out.terminated[ev] = (uint8_t)((f2 >> 8) & 1);
        if (a.out.truncated) a.out.truncated[ev] = (uint8_t)((f2 >> 9) & 1);
      }
    }
    if ((ph & RL_PHASE_COMMAND) || do_reset) {
      store_u8(sm, L.ishead, a.is_heading, env0, nvalid, ids, tid);
      store_u8(sm, L.isstand, a.is_standing, env0, nvalid, ids, tid);
    }
  }

  RL_STAMP(5);                  // stores issued
  // ---- ordered compaction of reset ids (ManagerBasedRLEnv.step: reset_buf.nonzero() [IL]) -----------
  if (compact) {
    if (nvalid == 0 && warp == kCompactWarp) compact_arrive();   // tiles past the end: an empty mask
    if (tid == kCompactWarp * 32) s_last = (early_prev == (unsigned)(vgrid - 1));
    tile_sync();
    if (s_last) {
      __threadfence();
      int* s_cnt = reinterpret_cast<int*>(sm);  // this CTA's tile is dead once its own stores have been issued
      if ((ph & RL_PHASE_OBS) && nvalid > 0) { if (tid == 0) bulk_wait_read0(); }
      tile_sync();
      // one mask per CTA; thread i takes mask base+i, a block-wide exclusive scan of the popcounts gives its
      // first output slot -> ids come out ascending
      const int G = vgrid;
      int run = 0;
#pragma unroll 1
      for (int base = 0; base < G; base += NT) {
        const int g = base + tid;
        unsigned m = (g < G) ? __ldcg(a.cta_mask + g) : 0u;
        const int cnt = __popc(m);
        int incl = cnt;   // inclusive scan inside the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (e >= d) incl += y; }
        if (e == 31) s_cnt[warp] = incl;
        tile_sync();
        if (warp == 0) {
          int wt = (e < NW) ? s_cnt[e] : 0, wi = wt;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, wi, d); if (e >= d) wi += y; }
          if (e < NW) s_cnt[e] = wi - wt;           // exclusive prefix of the warp totals
          if (e == 31) s_cnt[32] = wi;              // chunk total
        }
        tile_sync();
        int pos = run + s_cnt[warp] + incl - cnt;
        if (a.out.reset_ids)
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            a.out.reset_ids[pos++] = g * kE + b;
          }
        run += s_cnt[32];
        tile_sync();
      }
      if (tid == 0) {
        if (a.out.n_reset) *a.out.n_reset = run;
        *a.ticket = 0u;
      }
      return;
    }
  }
  if (do_reset && nvalid > 0) {
    const int n_cta = (n_total + kE - 1) / kE;
    if (tid == 0) s_last = (early_prev == (unsigned)(n_cta - 1));   // arrived right after the partials were written
    tile_sync();
    if (s_last) {
      __threadfence();
      // quantity q = lane, the CTAs' partials strided over the warps, then the warps' sums in warp order
      float* s_red = sm;   // this CTA's tile is dead once its own stores have been issued
      if ((ph & RL_PHASE_OBS) && tid == 0) bulk_wait_read0();
      tile_sync();
      for (int q = e; q < K + RL_MAX_DONE_TERMS + 2; q += 32) {
        float part = 0.f;
        for (int g = warp; g < n_cta; g += NW) part += __ldcg(a.log_partials + (size_t)g * RL_LOG_STRIDE + q);
        s_red[warp * RL_LOG_STRIDE + q] = part;
      }
      tile_sync();
      if (tid < K + RL_MAX_DONE_TERMS + 2) {
        float tot = 0.f;
        for (int w = 0; w < NW; ++w) tot += s_red[w * RL_LOG_STRIDE + tid];
        const RlResetLog& lg = a.out.reset_log;
        const int n_reset_total = reset_masked ? *a.out.n_reset : n_total;
        const float cnt = (float)max(n_reset_total, 1);
        if (n_reset_total == 0) tot = 0.f;
        if (tid < K) { if (lg.episode_sum_mean) lg.episode_sum_mean[tid] = tot / cnt; }
        else if (tid < K + RL_MAX_DONE_TERMS) { if (lg.done_term_count) lg.done_term_count[tid - K] = tot; }
        else if (lg.metric_mean) lg.metric_mean[tid - K - RL_MAX_DONE_TERMS] = tot / cnt;
      }
      if (tid == 0) *a.ticket = 0u;
    }
  }
  RL_STAMP(6);
  if ((ph & RL_PHASE_OBS) && nvalid > 0 && tid == 0) bulk_wait_read0();  // smem must outlive the bulk reads
  RL_STAMP(7);
#if RL_PERSISTENT
  }();
  tile_sync();   // every thread is done with the record (bulk reads drained by thread 0 above): the next tile may load
  }
#endif
}

// ---------------------------------------------------------------------------------------------------
// process_action: ActionManager.process_action + JointAction.process_actions [IL]
// ---------------------------------------------------------------------------------------------------
__global__ void process_action_kernel(int N, int slot, RlField new_action, RlField action, RlField prev_action,
                                      RlField target, RlField vel_target, unsigned long long* step_counter, int use_pdl) {
  if (use_pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (step_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *step_counter += 1ull;
  const RlActionCfg& ac = c_spec[slot].action;
  const int A = ac.n_actions;
  const int total = N * A;   // < 2^31, checked by the host
  const bool env_major = (new_action.env_stride == 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int env, col;
    if (env_major) { env = i % N; col = i / N; } else { col = i % A; env = i / A; }
    const float nv = static_cast<const float*>(new_action.ptr)[(long long)env * new_action.env_stride + col * new_action.comp_stride];
    float* ap = static_cast<float*>(action.ptr) + (long long)env * action.env_stride + col * action.comp_stride;
    if (prev_action.ptr)
      static_cast<float*>(prev_action.ptr)[(long long)env * prev_action.env_stride + col * prev_action.comp_stride] = *ap;
    *ap = nv;
    const RlField& dst = (ac.target_kind[col] == RL_ACTION_JOINT_VELOCITY) ? vel_target : target;
    if (dst.ptr) {
      float v = nv * ac.scale[col] + ac.offset[col];
      if (ac.has_clip) v = clampf(v, ac.clip_lo[col], ac.clip_hi[col]);
      static_cast<float*>(dst.ptr)[(long long)env * dst.env_stride + (long long)ac.joint_ids[col] * dst.comp_stride] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Reset events: reset_root_state_uniform (V/mdp/events.py:205-271) + reset_joints_by_scale [IL]. One thread per
// env being reset (they are few); isaaclab.utils.math [IL]: quat_from_euler_xyz, quat_mul, sample_uniform.
// ---------------------------------------------------------------------------------------------------
struct ResetStateArgs {
  int N, slot, has_ids;
  RlResetStateCfg cfg;
  RlField origins, pos, quat, lin, ang, jpos, jvel;
  const uint8_t* terminated;
  const uint8_t* truncated;
  const int32_t* env_ids;
  const int32_t* n_env_ids;
  RlRandom rnd;
  const float* uniforms;   // [12 + 2J][N] or NULL
  const uint8_t* pits;     // [N] is_env_assigned_to_terrain(env, "pits") or NULL
};
__global__ void reset_scene_state_kernel(const ResetStateArgs a) {
  const RlStepSpec& S = c_spec[a.slot];
  const int n = a.has_ids ? *a.n_env_ids : a.N;
  RandState rs;
  rs.seed = a.rnd.seed;
  rs.step = a.rnd.step + (a.rnd.step_counter ? *a.rnd.step_counter : 0ull);
  rs.env_id_offset = a.rnd.env_id_offset;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    long long env = i;
    if (a.has_ids) env = a.env_ids[i];
    else if (!((a.terminated && a.terminated[i]) || (a.truncated && a.truncated[i]))) continue;
    float u[12];
    if (a.uniforms) {
#pragma unroll
      for (int q = 0; q < 12; ++q) u[q] = a.uniforms[(long long)q * a.N + env];
    } else {
#pragma unroll
      for (int blk = 0; blk < 3; ++blk) {
        const uint4 r = rl_philox(rs, env, RL_STREAM_RESET_STATE, (uint32_t)blk);
        u[4 * blk] = u01(r.x); u[4 * blk + 1] = u01(r.y); u[4 * blk + 2] = u01(r.z); u[4 * blk + 3] = u01(r.w);
      }
    }
    const RlResetStateCfg& c = a.cfg;
    const float* org = static_cast<const float*>(a.origins.ptr);
    if (a.pits != nullptr && a.pits[env]) {
      // V/mdp/events.py:237-244: envs assigned to the "pits" sub-terrain get the default root state at their
      // origin, zero velocity, no random perturbation (the joints are still reset by reset_joints_by_scale below)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float o = org ? org[env * a.origins.env_stride + (long long)q * a.origins.comp_stride] : 0.f;
        st_f(a.pos, env, q, c.default_root_state[q] + o);
        st_f(a.lin, env, q, 0.f);
        st_f(a.ang, env, q, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) st_f(a.quat, env, q, c.default_root_state[3 + q]);
    } else {
    float pose[6], vel[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {   // sample_uniform [IL]: (hi - lo) * u + lo
      pose[q] = (c.pose_hi[q] - c.pose_lo[q]) * u[q] + c.pose_lo[q];
      vel[q] = (c.vel_hi[q] - c.vel_lo[q]) * u[6 + q] + c.vel_lo[q];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float o = org ? org[env * a.origins.env_stride + (long long)q * a.origins.comp_stride] : 0.f;
      st_f(a.pos, env, q, (c.default_root_state[q] + o) + pose[q]);
    }
    // quat_from_euler_xyz [IL] (roll, pitch, yaw)
    const float cr = rl_cosf(pose[3] * 0.5f), sr = rl_sinf(pose[3] * 0.5f);
    const float cp = rl_cosf(pose[4] * 0.5f), sp = rl_sinf(pose[4] * 0.5f);
    const float cy = rl_cosf(pose[5] * 0.5f), sy = rl_sinf(pose[5] * 0.5f);
    const float dw = (cy * cr) * cp + (sy * sr) * sp;
    const float dx = (cy * sr) * cp - (sy * cr) * sp;
    const float dy = (cy * cr) * sp + (sy * sr) * cp;
    const float dz = (sy * cr) * cp - (cy * sr) * sp;
    // quat_mul(default_quat, delta) [IL]
    const float w1 = c.default_root_state[3], x1 = c.default_root_state[4], y1 = c.default_root_state[5], z1 = c.default_root_state[6];
    const float ww = (z1 + x1) * (dx + dy), yy = (w1 - y1) * (dw + dz), zz = (w1 + y1) * (dw - dz);
    const float xx = ww + yy + zz;
    const float qq = 0.5f * (xx + (z1 - x1) * (dx - dy));
    st_f(a.quat, env, 0, qq - ww + (z1 - y1) * (dy - dz));
    st_f(a.quat, env, 1, qq - xx + (x1 + w1) * (dx + dw));
    st_f(a.quat, env, 2, qq - yy + (w1 - x1) * (dy + dz));
    st_f(a.quat, env, 3, qq - zz + (z1 + y1) * (dw - dx));
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      st_f(a.lin, env, q, c.default_root_state[7 + q] + vel[q]);
      st_f(a.ang, env, q, c.default_root_state[10 + q] + vel[3 + q]);
    }
    }
    // reset_joints_by_scale [IL]
    const int J = S.num_joints;
    for (int j = 0; j < J; ++j) {
      float up, uv;
      if (a.uniforms) {
        up = a.uniforms[(long long)(12 + j) * a.N + env];
        uv = a.uniforms[(long long)(12 + J + j) * a.N + env];
      } else {
        const uint4 r = rl_philox(rs, env, RL_STREAM_RESET_JOINTS, (uint32_t)(j >> 1));
        up = u01((j & 1) ? r.z : r.x); uv = u01((j & 1) ? r.w : r.y);
      }
      float qp = S.default_joint_pos[j] * ((c.joint_pos_scale_hi - c.joint_pos_scale_lo) * up + c.joint_pos_scale_lo);
      float qv = S.default_joint_vel[j] * ((c.joint_vel_scale_hi - c.joint_vel_scale_lo) * uv + c.joint_vel_scale_lo);
      qp = clampf(qp, S.soft_pos_limit_lo[j], S.soft_pos_limit_hi[j]);
      qv = clampf(qv, -S.soft_vel_limit[j], S.soft_vel_limit[j]);
      if (a.jpos.ptr) st_f(a.jpos, env, j, qp);
      if (a.jvel.ptr) st_f(a.jvel, env, j, qv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// ContactSensor update [IL]: history roll (or ring-slot write) + air / contact timers. One thread per
// (env, history element) for the history, one per (env, timer body) for the timers.
// ---------------------------------------------------------------------------------------------------
struct SensorArgs {
  int N, B, T, Bt, ring_slot;
  float dt, thr;
  RlField net, hist, cair, lair, ccon, lcon;
  int t2h[RL_MAX_TIME_BODIES];
};
__global__ void contact_sensor_kernel(const SensorArgs a) {
  // Work items: N * B*3 history elements, then N * Bt timer bodies. Consecutive threads walk the contiguous axis of
  // the tensor they write (components for AoS rows, envs for SoA) so that either layout coalesces.
  const int n_hist = a.N * a.B * 3, total = n_hist + a.N * a.Bt;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < n_hist) {
      int env, item;
      if (a.hist.comp_stride == 1) { item = i % (a.B * 3); env = i / (a.B * 3); } else { env = i % a.N; item = i / a.N; }
      const float v = static_cast<const float*>(a.net.ptr)[(long long)env * a.net.env_stride + (long long)item * a.net.comp_stride];
      float* h = static_cast<float*>(a.hist.ptr) + (long long)env * a.hist.env_stride;
      const long long cs = a.hist.comp_stride;
      if (a.ring_slot >= 0) {
        h[(long long)(a.ring_slot * a.B * 3 + item) * cs] = v;
      } else {
        // newest first: slot 0 <- net, slot t <- old slot t-1. All loads first, then all stores (hist_len <= 8):
        // one memory round trip instead of a load -> store chain per slot
        float old[8];
#pragma unroll
        for (int t = 0; t < 7; ++t)
          if (t < a.T - 1) old[t] = h[(long long)(t * a.B * 3 + item) * cs];
        h[(long long)item * cs] = v;
#pragma unroll
        for (int t = 1; t < 8; ++t)
          if (t < a.T) h[(long long)(t * a.B * 3 + item) * cs] = old[t - 1];
      }
    } else {
      const int j = i - n_hist;
      int env, f;
      if (a.cair.comp_stride == 1 && a.Bt > 1) { f = j % a.Bt; env = j / a.Bt; } else { env = j % a.N; f = j / a.N; }
      const int b = a.t2h[f];
      const float* net = static_cast<const float*>(a.net.ptr) + (long long)env * a.net.env_stride;
      const float fx = net[(long long)(3 * b) * a.net.comp_stride], fy = net[(long long)(3 * b + 1) * a.net.comp_stride],
                  fz = net[(long long)(3 * b + 2) * a.net.comp_stride];
      const bool contact = sqrtf((fx * fx + fy * fy) + fz * fz) > a.thr;
      float* ca = static_cast<float*>(a.cair.ptr) + (long long)env * a.cair.env_stride + (long long)f * a.cair.comp_stride;
      float* la = static_cast<float*>(a.lair.ptr) + (long long)env * a.lair.env_stride + (long long)f * a.lair.comp_stride;
      float* cc = static_cast<float*>(a.ccon.ptr) + (long long)env * a.ccon.env_stride + (long long)f * a.ccon.comp_stride;
      float* lc = static_cast<float*>(a.lcon.ptr) + (long long)env * a.lcon.env_stride + (long long)f * a.lcon.comp_stride;
      const float cur_air = *ca, cur_con = *cc;
      const bool first_contact = (cur_air > 0.f) && contact, first_detach = (cur_con > 0.f) && !contact;
      if (first_contact) *la = cur_air + a.dt;
      *ca = contact ? 0.f : cur_air + a.dt;
      if (first_detach) *lc = cur_con + a.dt;
      *cc = contact ? cur_con + a.dt : 0.f;
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
  int NW;                 // warps per tile (kE = 32 envs per tile is fixed)
  int tiles;              // tiles per CTA: 1, or 2 (baked specs at 16 warps whose two records fit one SM)
  Layout L;
  Schedule* sched_dev;
  Schedule sched;         // host copy of the schedule of the current launch config
  unsigned int* ticket;
  uint32_t* cta_mask;
  float* log_partials;
  int cta_mask_cap;
  RlRewardTerm* adhoc_dev;
  int sm_count;
  size_t smem_optin;      // largest dynamic shared memory a CTA may ask for on this device
  int use_pdl;
  long long* dbg;
  int baked;   // index into RL_BAKED_LIST when the spec equals a build-time specialised one, else -1
};

int rl_ctx_device_of(const RlCtx* ctx) { return ctx->device; }
const RlStepSpec* rl_ctx_spec_of(const RlCtx* ctx) { return &ctx->spec; }

namespace {


bool to_fd(const RlField& f, FieldD* out) {
  if (f.env_stride > 0x7fffffffLL || f.comp_stride > 0x7fffffffLL || f.env_stride < 0 || f.comp_stride < 0) return false;
  out->ptr = f.ptr; out->es = (int)f.env_stride; out->cs = (int)f.comp_stride;
  return true;
}
bool vec4_ok(const FieldD& d, int ncomp) {
  return d.ptr != nullptr && d.es == 1 && (ncomp == 1 || (d.cs % 4) == 0) && (reinterpret_cast<uintptr_t>(d.ptr) & 15u) == 0;
}

// Field descriptors + masks of one launch. `state` may be NULL (reset-only launches).
int fill_args(RlCtx* ctx, KArgs& a, int64_t num_envs, const RlStateView* st, const RlMdpState* mdp, const RlStepOut* out,
              const RlRandom* rnd, uint32_t ph, int mode) {
  const RlStepSpec& s = ctx->spec;
  memset(&a, 0, sizeof(a));
  a.N = (int)num_envs; a.slot = ctx->slot; a.phases = ph;
  a.L = ctx->L; a.sched = ctx->sched_dev;
  {
    const Schedule& sc = ctx->sched;
    for (int k = 0; k < s.num_reward_terms; ++k) {
      a.rw_weight[k] = s.rewards[k].weight;
      if (sc.late[k]) a.rw_late |= 1ull << k;
      if (s.rewards[k].type == RL_REW_IS_TERMINATED) a.rw_isterm |= 1ull << k;
      if (s.rewards[k].weight == 0.f) a.rw_zero |= 1ull << k;
    }
  }
  a.ticket = ctx->ticket; a.cta_mask = ctx->cta_mask; a.log_partials = ctx->log_partials; a.use_pdl = ctx->use_pdl;
  a.dbg = ctx->dbg;
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
  ok &= to_fd(S(&RlStateView::net_forces_w_history), &a.hist); ok &= to_fd(S(&RlStateView::ray_hits_z), &a.rays);
  ok &= to_fd(mdp->command, &a.in[IF_CMD]); ok &= to_fd(mdp->heading_target, &a.in[IF_HEAD]);
  ok &= to_fd(mdp->time_left, &a.in[IF_TLEFT]); ok &= to_fd(mdp->metric_error_vel_xy, &a.in[IF_MXY]);
  ok &= to_fd(mdp->metric_error_vel_yaw, &a.in[IF_MYAW]); ok &= to_fd(mdp->episode_length, &a.in[IF_EPLEN]);
  ok &= to_fd(mdp->episode_sums, &a.in[IF_SUMS]);
  ok &= to_fd(mdp->action, &a.in[IF_ACT]); ok &= to_fd(mdp->prev_action, &a.in[IF_PACT]);
  ok &= to_fd(mdp->is_heading_env, &a.is_heading); ok &= to_fd(mdp->is_standing_env, &a.is_standing);
  if (!ok) return fail(RL_EINVAL, "field strides must be non-negative and below 2^31 elements%s", "");
  a.in[IF_CMDU] = FieldD{a.rnd.cmd_uniforms, 1, (int)num_envs};
  const bool need_hist = (mode == 1) || (ph & (RL_PHASE_DONES | RL_PHASE_REWARDS));
  uint32_t m = (1u << IF_ROOT_POS) | (1u << IF_QUAT) | (1u << IF_LIN_VEL) | (1u << IF_ANG_VEL) | (1u << IF_JPOS) |
               (1u << IF_JVEL) | (1u << IF_CMD) | (1u << IF_EPLEN) | (1u << IF_ACT);
  if (need_hist)
    m |= (1u << IF_PACT) | (1u << IF_JACC) | (1u << IF_JTAU) | (1u << IF_CAIR) | (1u << IF_LAIR) | (1u << IF_CCON) | (1u << IF_LCON) |
         (1u << IF_BPOS) | (1u << IF_BVEL);
  if (mode == 0 && (ph & RL_PHASE_REWARDS)) m |= 1u << IF_SUMS;
  if (mode == 0 && (ph & RL_PHASE_COMMAND)) m |= (1u << IF_HEAD) | (1u << IF_TLEFT) | (1u << IF_MXY) | (1u << IF_MYAW) | (1u << IF_CMDU);
  if (mode == 0 && (ph & RL_PHASE_RESET)) m |= (1u << IF_SUMS) | (1u << IF_MXY) | (1u << IF_MYAW) | (1u << IF_HEAD) | (1u << IF_CMDU);
  if (mode == 0 && (ph & RL_PHASE_OBS) && s.num_rays > 0) m |= 1u << IF_RAYPOS;
  uint32_t v4 = 0;
  for (int f = 0; f < IF_COUNT; ++f) {
    if (a.in[f].ptr == nullptr) m &= ~(1u << f);
    else if (vec4_ok(a.in[f], in_field_ncomp(s, f))) v4 |= 1u << f;
  }
  a.in_mask = m; a.in_vec4 = v4;
  for (int f = 0; f < IF_COUNT; ++f) a.in[f].meta = in_field_ncomp(s, f) | (in_field_word(s, f) << 16);
  // outputs
  if (mode == 0) {
    a.outf[OF_REWARD] = FieldD{a.out.reward, 1, 0};
    a.outf[OF_EPLEN] = a.in[IF_EPLEN]; a.outf[OF_SUMS] = a.in[IF_SUMS];
    FieldD sr; if (!to_fd(a.out.step_reward, &sr)) return fail(RL_EINVAL, "bad step_reward strides%s", "");
    a.outf[OF_STEPR] = sr;
    a.outf[OF_CMD] = a.in[IF_CMD]; a.outf[OF_HEAD] = a.in[IF_HEAD]; a.outf[OF_TLEFT] = a.in[IF_TLEFT];
    a.outf[OF_MXY] = a.in[IF_MXY]; a.outf[OF_MYAW] = a.in[IF_MYAW];
    a.outf[OF_ACT] = a.in[IF_ACT]; a.outf[OF_PACT] = a.in[IF_PACT];
    uint32_t om = 0;
    if (ph & RL_PHASE_DONES) om |= 1u << OF_EPLEN;
    if (ph & RL_PHASE_REWARDS) om |= (1u << OF_REWARD) | (1u << OF_SUMS) | (1u << OF_STEPR);
    if (ph & (RL_PHASE_COMMAND | RL_PHASE_RESET)) om |= (1u << OF_CMD) | (1u << OF_HEAD) | (1u << OF_TLEFT) | (1u << OF_MXY) | (1u << OF_MYAW);
    if (ph & RL_PHASE_RESET) om |= (1u << OF_SUMS) | (1u << OF_EPLEN) | (1u << OF_ACT) | (1u << OF_PACT);
    const int oc[OF_COUNT] = {1, 1, s.num_reward_terms, s.num_reward_terms, 3, 1, 1, 1, 1, s.action.n_actions, s.action.n_actions};
    uint32_t ov4 = 0;
    for (int f = 0; f < OF_COUNT; ++f) {
      if (a.outf[f].ptr == nullptr) om &= ~(1u << f);
      else if (vec4_ok(a.outf[f], oc[f])) ov4 |= 1u << f;
    }
    a.out_mask = om; a.out_vec4 = ov4;
    for (int f = 0; f < OF_COUNT; ++f) a.outf[f].meta = out_field_ncomp(s, f) | (out_field_word(ctx->L, f) << 16);
  }
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

template <class P, int NW, int MODE, bool DBG, int TILES>
int launch_step_variant(RlCtx* ctx, const KArgs& a_in, int n_items, cudaStream_t st) {
  // every tile record starts 128-byte aligned (bulk copies need 16)
  const int tile_words = (ctx->L.total_words + 31) & ~31;
  const size_t smem = TILES > 1 ? (size_t)tile_words * 4 * TILES : (size_t)ctx->L.total_words * 4;
  static thread_local int configured_device = -1;
  static thread_local size_t configured_smem = 0;
  if (configured_device != ctx->device || configured_smem < smem) {
    CUDA_TRY(cudaFuncSetAttribute(mdp_step_kernel<P, NW, MODE, DBG, TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_device = ctx->device; configured_smem = smem;
  }
  const int vgrid = (n_items + kE - 1) / kE;
  if (vgrid <= 0) return RL_OK;
  KArgs a = a_in;
  a.vgrid = vgrid; a.tile_words = tile_words;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((vgrid + TILES - 1) / TILES); cfg.blockDim = dim3(NW * 32 * TILES); cfg.dynamicSmemBytes = smem; cfg.stream = st;
#if RL_PERSISTENT
  {  // no more CTAs than are resident at once: two per SM when two records (and 2 x 512 threads x 64 registers) fit
    const int per_sm = (TILES == 1 && NW * 32 <= 512 && 2 * (smem + 1024) <= (size_t)233472) ? 2 : 1;
    const int resident = ctx->sm_count * per_sm;
    if ((int)cfg.gridDim.x > resident) cfg.gridDim = dim3(resident);
  }
#endif
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = a.use_pdl ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, mdp_step_kernel<P, NW, MODE, DBG, TILES>, a));
  return RL_OK;
}
// the clock-stamp variant (rl_ctx_set_debug_buffer) is a separate instantiation: the production kernel carries
// no trace of it. Two tiles per CTA (rl_ctx_set_launch_config(ctx, 64, 16)) exist for the build-time specialised
// kernels at 16 warps per tile only - 1024 threads, two records in shared memory.
template <class P, int NW, int MODE>
int launch_step(RlCtx* ctx, const KArgs& a, int n_items, cudaStream_t st) {
  if constexpr (MODE == 0) {
    if (a.dbg != nullptr) return launch_step_variant<P, NW, MODE, true, 1>(ctx, a, n_items, st);
    if constexpr (P::kStatic && NW == 16) {
      if (ctx->tiles == 2) return launch_step_variant<P, NW, MODE, false, 2>(ctx, a, n_items, st);
    }
  }
  return launch_step_variant<P, NW, MODE, false, 1>(ctx, a, n_items, st);
}

// warps per CTA compiled for the generic kernel / for every baked spec
#define RL_DYN_CONFIGS(X) X(4) X(8) X(16)
#define RL_STATIC_CONFIGS(X) X(4) X(8) X(16)

template <class P, int MODE>
int dispatch_config(RlCtx* ctx, const KArgs& a, int n_items, cudaStream_t st, bool* found) {
  *found = true;
  if constexpr (P::kStatic) {
#define RL_CASE(W_) if (ctx->NW == (W_)) return launch_step<P, W_, MODE>(ctx, a, n_items, st);
    RL_STATIC_CONFIGS(RL_CASE)
#undef RL_CASE
  } else {
#define RL_CASE(W_) if (ctx->NW == (W_)) return launch_step<P, W_, MODE>(ctx, a, n_items, st);
    RL_DYN_CONFIGS(RL_CASE)
#undef RL_CASE
  }
  *found = false;
  return RL_OK;
}

template <int MODE>
int dispatch_step(RlCtx* ctx, const KArgs& a, int n_items, cudaStream_t st) {
  bool found = false;
  if (MODE == 0 && ctx->baked >= 0) {
    int idx = 0, rc = RL_OK;
#define RL_TRY_BAKED(B) if (idx++ == ctx->baked) rc = dispatch_config<StaticPolicy<baked::B>, 0>(ctx, a, n_items, st, &found);
    RL_BAKED_LIST(RL_TRY_BAKED)
#undef RL_TRY_BAKED
    if (found) return rc;
  }
  // generic kernel (compiled for 4, 8 and 16 warps). The single-term mode ignores the schedule, so any
  // configured warp count maps to a compiled one; the step mode needs the schedule's warp count exactly.
  RlCtx tmp = *ctx;
  if (MODE == 1) tmp.NW = tmp.NW >= 16 ? 16 : (tmp.NW >= 8 ? 8 : 4);
  int rc = dispatch_config<DynPolicy, MODE>(&tmp, a, n_items, st, &found);
  if (!found) return fail(RL_EINVAL, "the generic kernel is compiled for 4, 8 or 16 warps per CTA, not %s%lld", "", ctx->NW);
  return rc;
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

bool g_slots[16][RL_SPEC_SLOTS];

// two tiles per CTA: a build-time specialised kernel at 16 warps per tile, both records resident in one CTA
bool two_tiles_ok(const RlCtx* ctx) {
  const size_t tile_bytes = (size_t)((ctx->L.total_words + 31) & ~31) * 4;
  return ctx->baked >= 0 && ctx->NW == 16 && 2 * tile_bytes <= ctx->smem_optin;
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
  RL_SZ(RlStepOut); RL_SZ(RlRandom); RL_SZ(RlResetLog); RL_SZ(RlResetStateCfg);
  RL_SZ(RlActuatorCfg); RL_SZ(RlTerrainGrid); RL_SZ(RlHeightField);
#undef RL_SZ
  return -1;
}

int64_t rl_tile_record_bytes(const RlStepSpec* spec) {
  if (!spec || validate_spec(spec) != RL_OK) return -1;
  return (int64_t)make_layout(*spec).total_words * 4;
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
  ctx->NW = 16;
  ctx->L = make_layout(ctx->spec);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  ctx->tiles = 1;
  if ((size_t)ctx->L.total_words * 4 > (size_t)prop.sharedMemPerBlockOptin) {
    delete ctx;
    return fail(RL_EUNSUPPORTED, "the tile of %s%lld envs needs %lld bytes of shared memory", "", kE, (long long)make_layout(*spec).total_words * 4);
  }
  CUDA_TRY(cudaMalloc(&ctx->sched_dev, sizeof(Schedule)));
  ctx->sched = make_schedule(ctx->spec, ctx->NW);
  CUDA_TRY(cudaMemcpy(ctx->sched_dev, &ctx->sched, sizeof(Schedule), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpyToSymbol(c_spec, spec, sizeof(RlStepSpec), sizeof(RlStepSpec) * slot));
  CUDA_TRY(cudaMalloc(&ctx->ticket, sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(ctx->ticket, 0, sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc(&ctx->adhoc_dev, sizeof(RlRewardTerm) * 64));
  rc = ensure_scratch(ctx, 4096);
  if (rc != RL_OK) return rc;
  ctx->baked = -1;
  {
    int idx = 0;
#define RL_MATCH_BAKED(B) if (ctx->baked < 0 && memcmp(spec, &baked::B::spec, sizeof(RlStepSpec)) == 0) ctx->baked = idx; ++idx;
    RL_BAKED_LIST(RL_MATCH_BAKED)
#undef RL_MATCH_BAKED
    // RL_MDPSTEP_GENERIC=1 forces the generic (table-driven) kernel: A/B measurements and tests of that path
    const char* force_generic = getenv("RL_MDPSTEP_GENERIC");
    if (force_generic && force_generic[0] == '1') ctx->baked = -1;
    // RL_MDPSTEP_TILES=2: two tiles per CTA wherever that configuration exists (A/B measurements, test runs of
    // that path); contexts it does not apply to keep one tile
    const char* want_tiles = getenv("RL_MDPSTEP_TILES");
    if (want_tiles && want_tiles[0] == '2' && two_tiles_ok(ctx)) ctx->tiles = 2;
  }
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
  if (ctx->sched_dev) cudaFree(ctx->sched_dev);
  g_slots[ctx->device][ctx->slot] = false;
  delete ctx;
}

int rl_ctx_set_launch_config(RlCtx* ctx, int envs_per_cta, int warps_per_cta) {
  if (!ctx) return fail(RL_EINVAL, "null ctx%s", "");
  if (envs_per_cta != 0 && envs_per_cta != kE && envs_per_cta != 2 * kE)
    return fail(RL_EINVAL, "envs_per_cta must be %s%lld (one tile, one lane per env) or twice that (two tiles per CTA)", "", kE);
  const int nw = warps_per_cta > 0 ? warps_per_cta : 16;
  if (nw != 4 && nw != 8 && nw != 16) return fail(RL_EINVAL, "warps_per_cta must be 4, 8 or 16%s, got %lld", "", nw);
  if (envs_per_cta == 2 * kE) {
    RlCtx probe = *ctx;
    probe.NW = nw;
    if (!two_tiles_ok(&probe))
      return fail(RL_EUNSUPPORTED, "two tiles per CTA need a build-time specialised spec, 16 warps per tile and 2 x %s%lld bytes of shared memory", "",
                  (long long)ctx->L.total_words * 4);
  }
  DeviceGuard guard(ctx->device);
  ctx->sched = make_schedule(ctx->spec, nw);
  CUDA_TRY(cudaMemcpy(ctx->sched_dev, &ctx->sched, sizeof(Schedule), cudaMemcpyHostToDevice));  // synchronous: not for hot loops
  ctx->NW = nw;
  ctx->tiles = envs_per_cta == 2 * kE ? 2 : 1;
  return RL_OK;
}

int rl_ctx_get_launch_config(const RlCtx* ctx, int* envs_per_cta, int* warps_per_tile) {
  if (!ctx || !envs_per_cta || !warps_per_tile) return fail(RL_EINVAL, "null argument%s", "");
  *envs_per_cta = kE * ctx->tiles;
  *warps_per_tile = ctx->NW;
  return RL_OK;
}

int rl_ctx_get_schedule(RlCtx* ctx, int32_t* out, int32_t* n_tasks) {
  if (!ctx || !out || !n_tasks) return fail(RL_EINVAL, "null argument%s", "");
  const Schedule& sc = ctx->sched;
  *n_tasks = sc.n;
  for (int i = 0; i < sc.n; ++i) {
    const Task& t = sc.t[i];
    const int32_t row[8] = {t.kind, t.a, t.b, t.owner, t.lo, t.hi, t.col0, t.pad};
    for (int q = 0; q < 8; ++q) out[i * 8 + q] = row[q];
  }
  return RL_OK;
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

int rl_reset_scene_state(RlCtx* ctx, int64_t num_envs, const RlResetStateCfg* cfg, const RlField* env_origins,
                         const RlStateView* state, const uint8_t* terminated, const uint8_t* truncated,
                         const int32_t* env_ids, const int32_t* n_env_ids, const RlRandom* rnd,
                         const float* uniforms, const uint8_t* assigned_to_pits, void* stream) {
  if (!ctx || !cfg || !state || !rnd) return fail(RL_EINVAL, "rl_reset_scene_state: null argument%s", "");
  if (num_envs <= 0) return RL_OK;
  if (!state->root_pos_w.ptr || !state->root_quat_w.ptr || !state->root_lin_vel_w.ptr || !state->root_ang_vel_w.ptr)
    return fail(RL_EINVAL, "rl_reset_scene_state: the four root state fields are required%s", "");
  if (env_ids && !n_env_ids) return fail(RL_EINVAL, "rl_reset_scene_state: env_ids needs n_env_ids%s", "");
  if (!env_ids && !terminated && !truncated) return fail(RL_EINVAL, "rl_reset_scene_state: env_ids or the done masks are required%s", "");
  ResetStateArgs a;
  memset(&a, 0, sizeof(a));
  a.N = (int)num_envs; a.slot = ctx->slot; a.has_ids = env_ids != nullptr;
  a.cfg = *cfg;
  if (env_origins) a.origins = *env_origins;
  a.pos = state->root_pos_w; a.quat = state->root_quat_w; a.lin = state->root_lin_vel_w; a.ang = state->root_ang_vel_w;
  a.jpos = state->joint_pos; a.jvel = state->joint_vel;
  a.terminated = terminated; a.truncated = truncated; a.env_ids = env_ids; a.n_env_ids = n_env_ids;
  a.rnd = *rnd; a.uniforms = uniforms; a.pits = assigned_to_pits;
  DeviceGuard guard(ctx->device);
  const int threads = 128;
  const int blocks = (int)((num_envs + threads - 1) / threads);
  reset_scene_state_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return RL_OK;
}

int rl_contact_sensor_update(RlCtx* ctx, int64_t num_envs, const RlField* net_forces_w, const RlStateView* state,
                             const int32_t* time_body_to_hist, float dt, float force_threshold, int32_t ring_slot,
                             void* stream) {
  if (!ctx || !net_forces_w || !state) return fail(RL_EINVAL, "rl_contact_sensor_update: null argument%s", "");
  if (num_envs <= 0) return RL_OK;
  const RlStepSpec& s = ctx->spec;
  if (!net_forces_w->ptr || !state->net_forces_w_history.ptr) return fail(RL_EINVAL, "rl_contact_sensor_update: net_forces_w and net_forces_w_history required%s", "");
  if (s.num_time_bodies > 0 && (!time_body_to_hist || !state->current_air_time.ptr || !state->last_air_time.ptr ||
                                !state->current_contact_time.ptr || !state->last_contact_time.ptr))
    return fail(RL_EINVAL, "rl_contact_sensor_update: the four timer fields and time_body_to_hist are required%s", "");
  if (ring_slot >= s.hist_len) return fail(RL_EINVAL, "rl_contact_sensor_update: ring_slot %s%lld outside the history length", "", ring_slot);
  if (!(dt > 0.f)) return fail(RL_EINVAL, "rl_contact_sensor_update: dt must be positive%s", "");
  SensorArgs a;
  memset(&a, 0, sizeof(a));
  a.N = (int)num_envs; a.B = s.num_hist_bodies; a.T = s.hist_len; a.Bt = s.num_time_bodies; a.ring_slot = ring_slot;
  a.dt = dt; a.thr = force_threshold;
  a.net = *net_forces_w; a.hist = state->net_forces_w_history;
  a.cair = state->current_air_time; a.lair = state->last_air_time;
  a.ccon = state->current_contact_time; a.lcon = state->last_contact_time;
  for (int f = 0; f < s.num_time_bodies; ++f) {
    if (time_body_to_hist[f] < 0 || time_body_to_hist[f] >= s.num_hist_bodies)
      return fail(RL_EINVAL, "rl_contact_sensor_update: time_body_to_hist[%s%lld] outside the history bodies", "", f);
    a.t2h[f] = time_body_to_hist[f];
  }
  const long long total = num_envs * (long long)(a.B * 3 + a.Bt);
  if (total >= (1ll << 31)) return fail(RL_EINVAL, "rl_contact_sensor_update: problem too large for 32-bit indexing%s", "");
  if (total == 0) return RL_OK;
  DeviceGuard guard(ctx->device);
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads);
  contact_sensor_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return RL_OK;
}

int rl_process_action(RlCtx* ctx, int64_t num_envs, const RlField* new_action, const RlMdpState* mdp,
                      const RlField* joint_target, const RlField* joint_vel_target, uint64_t* step_counter, void* stream) {
  if (!ctx || !new_action || !mdp) return fail(RL_EINVAL, "rl_process_action: null argument%s", "");
  if (num_envs <= 0) return RL_OK;
  if (!new_action->ptr || !mdp->action.ptr) return fail(RL_EINVAL, "rl_process_action: action pointers required%s", "");
  DeviceGuard guard(ctx->device);
  RlField tgt = joint_target ? *joint_target : RlField{nullptr, 0, 0};
  RlField vtgt = joint_vel_target ? *joint_vel_target : RlField{nullptr, 0, 0};
  const long long total = num_envs * ctx->spec.action.n_actions;
  if (total >= (1ll << 31)) return fail(RL_EINVAL, "rl_process_action: num_envs * n_actions must stay below 2^31%s", "");
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
                              mdp->prev_action, tgt, vtgt, (unsigned long long*)step_counter, ctx->use_pdl));
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
  if ((phases & RL_PHASE_RESET) && !env_ids && (!out->terminated || !out->truncated || !out->n_reset))
    return fail(RL_EINVAL, "rl_step: RL_PHASE_RESET without env_ids resets the envs flagged in out->terminated | out->truncated and needs out->n_reset%s", "");
  if ((phases & RL_PHASE_RESET) && (phases & (RL_PHASE_DONES | RL_PHASE_REWARDS))) return fail(RL_EINVAL, "rl_step: RESET combines with COMMAND/OBS only%s", "");
  if ((phases & RL_PHASE_RESET) && (!mdp->episode_sums.ptr || !mdp->prev_action.ptr || !mdp->heading_target.ptr || !mdp->time_left.ptr ||
      !mdp->is_heading_env.ptr || !mdp->is_standing_env.ptr || !mdp->metric_error_vel_xy.ptr || !mdp->metric_error_vel_yaw.ptr))
    return fail(RL_EINVAL, "rl_step: RL_PHASE_RESET needs every RlMdpState field%s", "");
  if (env_ids && (phases & RL_PHASE_COMPACT)) return fail(RL_EINVAL, "rl_step: compaction is not available on an env_ids subset%s", "");
  if ((phases & RL_PHASE_COMPACT) && !(phases & RL_PHASE_DONES)) return fail(RL_EINVAL, "rl_step: COMPACT needs DONES%s", "");
  DeviceGuard guard(ctx->device);
  const int grid = (int)((num_envs + kE - 1) / kE);
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
  int frc = fill_args(ctx, a, num_envs, state, mdp, out, rnd, phases, 0);
  if (frc != RL_OK) return frc;
  a.has_ids = env_ids != nullptr; a.env_ids = env_ids; a.n_env_ids = n_env_ids;
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
  const int grid = (int)((num_envs + kE - 1) / kE);
  if (grid > ctx->cta_mask_cap) {
    cudaStreamCaptureStatus cs;
    CUDA_TRY(cudaStreamIsCapturing((cudaStream_t)stream, &cs));
    if (cs != cudaStreamCaptureStatusNone) return fail(RL_EINVAL, "rl_reset_envs: first call at this num_envs must happen outside stream capture%s", "");
    int rc = ensure_scratch(ctx, grid);
    if (rc != RL_OK) return rc;
  }
  KArgs a;
  RlStepOut o;
  memset(&o, 0, sizeof(o));
  o.done_bits = const_cast<uint8_t*>(done_bits);
  if (log) o.reset_log = *log;
  int frc = fill_args(ctx, a, num_envs, nullptr, mdp, &o, rnd, RL_PHASE_RESET, 0);
  if (frc != RL_OK) return frc;
  a.has_ids = 1; a.env_ids = env_ids; a.n_env_ids = n_env_ids;
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
  int frc = fill_args(ctx, a, num_envs, state, mdp, nullptr, nullptr, 0, 1);
  if (frc != RL_OK) return frc;
  a.adhoc = dev; a.ext_terminated = terminated; a.term_out = out; a.use_pdl = 0;
  return dispatch_step<1>(ctx, a, (int)num_envs, st);
}

}  // extern "C"
