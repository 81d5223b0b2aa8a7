// mdp_terms.cuh - device code shared by the step kernels of libmdpstep.so (csrc/mdp_step.cu: the general kernel,
// csrc/mdp_step_v2.cu: the cluster kernels of the two launch kinds of an env step): spec access policies, the
// shared-memory record layout, the work schedule, PTX helpers and the term functions themselves - every reward term,
// the command update, the observation terms. Reference behaviour restated (paths relative to /root/reference, V/ =
// source/robot_lab/robot_lab/tasks/manager_based/locomotion/velocity/): V/mdp/rewards.py:22-687,
// V/mdp/observations.py:17-35, V/mdp/commands.py:22-85, V/velocity_env_cfg.py:106-254,379-664 and the IsaacLab
// manager loops / upstream terms listed in SURVEY.md Appendix A. Everything lives in an anonymous namespace: each
// translation unit gets its own copy (and its own __constant__ spec slots, filled by rl_ctx_create).
#ifndef RL_MDP_TERMS_CUH_
#define RL_MDP_TERMS_CUH_

#include "rl_common.cuh"

#include <new>
#include <utility>

#include "generated/baked_specs.cuh"

#define RL_SPEC_SLOTS 3
#define RL_PI_F 3.14159265358979323846f
#define RL_LOG_STRIDE 64  // >= RL_MAX_REWARD_TERMS + RL_MAX_DONE_TERMS + 2

static __constant__ RlStepSpec c_spec[RL_SPEC_SLOTS];

// plain data types + constexpr helpers: a named namespace, so that the two translation units that include this header
// agree on the types they hand each other (RlCtx, KArgs)
namespace rlk {

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
// Loops inside the term functions (over joints, feet, bodies). The general kernel keeps them rolled: every SM executes all
// of its code once per launch, and code size is what that costs. The kernels of mdp_step_v2.cu define RL_TERM_LOOP as a
// full unroll before including this header: with a baked spec the trip counts, masks and addresses are constants, the
// independent parts of the iterations overlap, and the sums keep their order (bit-identical).
#ifndef RL_TERM_LOOP
#define RL_TERM_LOOP _Pragma("unroll 1")
#endif
// The SHORT loops - over the (four) feet, over the bodies of a contact mask. The kernels of mdp_step_v2.cu unroll these:
// measured per task (tools/v2_timeline.py, profiles/r2_summary.md) the feet terms were the longest tasks of a tile - the
// critical path of a 4096-env launch - because every iteration waited for a run-time indexed constant load, a call of the
// division subroutine and the previous iteration's quaternion chain; unrolled against the baked spec the indices are
// immediates, x / 2 and x / 4 are exact multiplications, and the feet's independent chains overlap. The order of every
// sum is unchanged (bit-identical results).
#ifndef RL_FEW_LOOP
#define RL_FEW_LOOP _Pragma("unroll 1")
#endif
constexpr int kE = 32;  // envs per CTA = lanes per warp: in the compute phase lane e of every warp owns env e
// (Round 1 cut the one long body sum - undesired_contacts over 15 bodies - into two parts that met through a lock-free
// arrival counter in shared memory. Round 2 removed it: compute-sanitizer racecheck flags the protocol, the parts rounded
// differently from the reference's single sum unless gated once, and the kernels that matter for speed - mdp_step_v2.cu -
// cache the contact norms instead, which makes the term short.)
// (Round-1 build variants RL_SHARED_NORMS / RL_SHARED_CTX / RL_PERSISTENT were measured in round 2 - profiles/
// r2_variant_probe.txt: norms + context prepass -9 % on the pre-reset launch, persistent tile loop < 3 % at any size - and
// removed; the cluster kernels of csrc/mdp_step_v2.cu carry the norm prepass.)

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
  int E;                                          // envs of the record (32 per tile x tiles of the CTA); SoA word w of env e = sm[w*E + e]
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
  int termv;                                     // [K] weighted value of every term (the manager-order sum reads them)
  int hnorm;                                     // [B] max over the history of |F_b| (written by a prepass; cached-norm kernels only)
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

__host__ __device__ constexpr Layout make_layout(const RlStepSpec& s, const int E = kE) {
  Layout L{};
  L.E = E;
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
  L.termv = take(K) * E;
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
// and one per observation term (the height scan in column chunks),
// balanced over the warps by a longest-processing-time greedy on rough instruction costs. constexpr, so a baked
// spec gets its schedule at compile time and every warp's code is straight-line.
// ---------------------------------------------------------------------------------------------------
enum { TK_REWARD = 0, TK_OBS = 1, TK_DONES = 2, TK_COMMAND = 3, TK_LOG = 4 /* reset logging + zeroing (mdp_step_v2.cu) */ };

struct Task {
  uint8_t kind, a, b, owner;   // REWARD: a = term k, b = half (0/1); OBS: a = group, b = term index
  uint16_t lo, hi;             // REWARD: body-index range [lo, hi); OBS: column range within the term
  uint16_t col0, pad;          // OBS: first column of the term inside the group row
};
struct Schedule {
  int n;
  Task t[RL_MAX_TASKS];
  uint8_t late[RL_MAX_REWARD_TERMS];    // term is finished in stage 2 (is_terminated)
};

__host__ __device__ constexpr int popc64(uint64_t m) { int n = 0; while (m) { m &= m - 1; ++n; } return n; }

__host__ __device__ constexpr int reward_cost(const RlRewardTerm& t, const RlStepSpec& s, int nbodies, const bool cached = false) {
  const int J = popc64(t.joint_mask), F = t.n_idx, T = s.hist_len;
  switch (t.type) {
    case RL_REW_JOINT_TORQUES_L2: case RL_REW_JOINT_VEL_L2: case RL_REW_JOINT_ACC_L2: case RL_REW_JOINT_DEVIATION_L1:
    case RL_REW_JOINT_POWER: case RL_REW_STAND_STILL: return 30 + 5 * J;
    case RL_REW_JOINT_POS_LIMITS: case RL_REW_JOINT_VEL_LIMITS: case RL_REW_JOINT_POS_PENALTY: return 40 + 8 * J;
    case RL_REW_JOINT_MIRROR: case RL_REW_ACTION_MIRROR: return 30 + 8 * F;
    case RL_REW_ACTION_SYNC: return 40 + 30 * F;
    case RL_REW_ACTION_RATE_L2: return 30 + 5 * s.action.n_actions;
    case RL_REW_UNDESIRED_CONTACTS: case RL_REW_CONTACT_FORCES: return cached ? 30 + nbodies * 5 : 30 + nbodies * T * 18;
    case RL_REW_TRACK_LIN_VEL_XY_EXP: case RL_REW_TRACK_ANG_VEL_Z_EXP: case RL_REW_TRACK_ANG_VEL_Z_WORLD_EXP: return 70;
    case RL_REW_TRACK_LIN_VEL_XY_YAW_FRAME_EXP: return 260;
    case RL_REW_FEET_AIR_TIME: case RL_REW_FEET_CONTACT: case RL_REW_FEET_CONTACT_WITHOUT_CMD: return 30 + 10 * F;
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: return 40 + 12 * F;
    case RL_REW_FEET_AIR_TIME_VARIANCE: return 40 + 40 * F;
    case RL_REW_FEET_GAIT: return 260;
    case RL_REW_FEET_STUMBLE: return 30 + 20 * F;
    case RL_REW_FEET_SLIDE: return cached ? 30 + F * 62 : 30 + F * (60 + T * 18);
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
    const bool body_sum = (t.type == RL_REW_UNDESIRED_CONTACTS || t.type == RL_REW_CONTACT_FORCES);
    sc.t[n] = Task{TK_REWARD, (uint8_t)k, 0, 0, 0, 64, 1, 0};
    cost[n++] = 30 + reward_cost(t, s, body_sum ? popc64(t.body_mask) : t.n_idx);
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
  const float* cj;        // per-joint constants [5][J] in device memory (q0, qd0, soft lo, soft hi, vel limit): one coalesced read
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

}  // namespace rlk

namespace {
using namespace rlk;


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
// The ONE thread that drew the last ticket of a launch: an acquiring read of the ticket word synchronises with every
// tile's release-arrival (they form one chain of read-modify-writes); the CTA barrier that follows extends that to the
// CTA's other threads, whose tail loads go to L2 (ld.cg). No fence: a fence would also wait for this CTA's own
// outstanding stores, with every other SM idle (profiles/r2_summary.md: 2 us of a 4096-env launch).
__device__ __forceinline__ void ticket_acquire(const unsigned int* ticket) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ticket) : "memory");
  (void)v;
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
// IEEE division, one code copy. A zero numerator (a reward term that is 0 for the env: most contact / air-time terms on most
// steps) sends the WHOLE warp through the ~30-instruction special-operand path of the division subroutine - 11 % of the
// pre-reset launch's instructions at 65536 envs (profiles/r2_summary.md). (+-0) / y for a finite normal y is (+-0) * y
// exactly (sign = xor of the signs), so those lanes divide 1 by y instead and take the product.
__device__ __noinline__ float rl_div(float x, float y) {
  const bool z = (x == 0.f) && (fabsf(y) >= 1.17549435e-38f) && (fabsf(y) <= 3.402823466e38f);
  const float q = (z ? 1.f : x) / y;
  return z ? x * y : q;
}
// x / (float)n for a small positive count n: halves and quarters are exact multiplications (the correctly rounded product
// and the correctly rounded quotient of the same real number), everything else goes through the division subroutine
__device__ __forceinline__ float div_count(float x, int n) {
  if (n == 1) return x;
  if (n == 2) return x * 0.5f;
  if (n == 4) return x * 0.25f;
  return rl_div(x, (float)n);
}
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
#define SMF(off, c) sm[(off) + (c) * L.E + e]
#define CJ(k, j) sm[L.cj + (k) * L.J + (j)]

__device__ __forceinline__ bool first_contact(const float* sm, const Layout& L, const Scalars& S, int e, int b) {
  const float t = SMF(L.ccon, b);
  return (t > 0.f) && (t < (S.step_dt + S.contact_time_abs_tol));
}
__device__ __forceinline__ V3 body_vec(const float* sm, const Layout& L, int off, int e, int b) {
  return V3{SMF(off, 3 * b + 0), SMF(off, 3 * b + 1), SMF(off, 3 * b + 2)};
}
// what the norm consumers call: the record's cached value (prepass) or the computation itself
// (CN: a compile-time bool in scope at the point of use - cached norms or not)
#define HIST_MAX_NORM(h, b) (CN ? SMF(L.hnorm, (b)) : hist_max_norm(h, S.hist_len, S.num_hist_bodies, b))   /* arguments are plain identifiers / subscripts */
// max over the history of |F_b| (net_forces_w_history[:, :, b].norm(-1).max(1))
__device__ __noinline__ float hist_max_norm(const float* h, int T, int B, int b) {
  // ONE copy for every term that uses it (undesired_contacts, contact_forces, feet_slide, feet_stumble, the
  // illegal-contact termination): warps in different terms keep the same few instruction lines hot instead of
  // evicting each other's private copies. A force of exactly 0 skips the IEEE sqrt (its special-case path).
  // max_t sqrt(ss_t) == sqrt(max_t ss_t) bit for bit: the correctly rounded square root is monotonic - ONE IEEE sqrt per
  // body instead of one per history sample
  float m2 = 0.f;
  RL_TERM_LOOP
  for (int t = 0; t < T; ++t) {
    const float* f = h + (t * B + b) * 3;
    const float ss = (f[0] * f[0] + f[1] * f[1]) + f[2] * f[2];
    m2 = (t == 0) ? ss : fmaxf(m2, ss);
  }
  const bool z = (m2 == 0.f);   // a body without contact: keep the warp off sqrt's special-operand path
  const float r = sqrtf(z ? 1.f : m2);
  return z ? 0.f : r;
}

// One reward term for env e: raw value (no weight, no dt). [lo, hi) restricts body-mask terms to a body-index
// range (callers pass 0, 64: the whole mask).
// `t` may be a build-time constant (scalar members fold into immediates); `tc` is the same term in __constant__
// memory and serves every run-time indexed list (a baked object indexed at run time would be a global-memory load).
template <bool CN>
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
      RL_TERM_LOOP
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) { const float v = SMF(off, j); s += v * v; }
      return s;
    }
    case RL_REW_JOINT_DEVIATION_L1: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jpos, j) - CJ(0, j));
      return s;
    }
    case RL_REW_JOINT_POS_LIMITS: {
      float s = 0.f;
      RL_TERM_LOOP
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
      RL_TERM_LOOP
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += clampf(fabsf(SMF(L.jvel, j)) - CJ(4, j) * t.p[0], 0.f, 1.f);
      return s;
    }
    case RL_REW_JOINT_POWER: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jvel, j) * SMF(L.jtau, j));
      return s;
    }
    case RL_REW_STAND_STILL: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) s += fabsf(SMF(L.jpos, j) - CJ(0, j));
      s *= (c.cmd_norm < t.p[0]) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_JOINT_POS_PENALTY: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int j = 0; j < J; ++j)
        if ((t.joint_mask >> j) & 1ull) { const float d = SMF(L.jpos, j) - CJ(0, j); s += d * d; }
      const float running = sqrtf(s);
      const bool moving = (c.cmd_norm > t.p[2]) || (c.vxy_norm > t.p[1]);
      return (moving ? running : t.p[0] * running) * c.gate;
    }
    case RL_REW_JOINT_MIRROR: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int i = 0; i < t.n_idx; ++i) {
        const float d = SMF(L.jpos, tc.idx_a[i]) - SMF(L.jpos, tc.idx_b[i]);
        s += d * d;
      }
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_MIRROR: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int i = 0; i < t.n_idx; ++i) {
        const float d = fabsf(SMF(L.act, tc.idx_a[i])) - fabsf(SMF(L.act, tc.idx_b[i]));
        s += d * d;
      }
      return (s * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_SYNC: {
      float r = 0.f;
      RL_TERM_LOOP
      for (int g = 0; g < t.n_idx; ++g) {
        const int start = tc.idx_b[g], n = tc.idx_c[g];
        if (n < 2) continue;
        float m = 0.f;
        for (int i = 0; i < n; ++i) m += fabsf(SMF(L.act, tc.idx_a[start + i]));
        m = div_count(m, n);
        float v = 0.f;
        for (int i = 0; i < n; ++i) { const float d = fabsf(SMF(L.act, tc.idx_a[start + i])) - m; v += d * d; }
        r += div_count(v, n);
      }
      return (r * t.p[0]) * c.gate;
    }
    case RL_REW_ACTION_RATE_L2: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int a = 0; a < S.n_actions; ++a) { const float d = SMF(L.act, a) - SMF(L.pact, a); s += d * d; }
      return s;
    }
    case RL_REW_UNDESIRED_CONTACTS: {
      float s = 0.f;
      RL_FEW_LOOP
      for (int b = lo; b < S.num_hist_bodies && b < hi; ++b)
        if (((t.body_mask >> b) & 1ull) && (HIST_MAX_NORM(h, b) > t.p[0])) s += 1.f;
      // one part of a split term (a restricted body range) returns its raw COUNT: the finisher adds the counts (exact)
      // and applies the gate once, as the reference does (V/mdp/rewards.py:672-675) - count_a * gate + count_b * gate
      // rounds differently from (count_a + count_b) * gate
      return (lo > 0 || hi < 64) ? s : s * c.gate;
    }
    case RL_REW_CONTACT_FORCES: {
      float s = 0.f;
      RL_FEW_LOOP
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
      RL_FEW_LOOP
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = tc.idx_a[i];
        s += (SMF(L.lair, b) - t.p[0]) * (first_contact(sm, L, S, e, b) ? 1.f : 0.f);
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_AIR_TIME_POSITIVE_BIPED: {
      int n_contact = 0;
      RL_FEW_LOOP
      for (int i = 0; i < t.n_idx; ++i) n_contact += (SMF(L.ccon, tc.idx_a[i]) > 0.f) ? 1 : 0;
      const bool single = (n_contact == 1);
      float r = INFINITY;
      RL_FEW_LOOP
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
      RL_FEW_LOOP
      for (int which = 0; which < 2; ++which) {
        const int off = which == 0 ? L.lair : L.lcon;
        float mean = 0.f, m2 = 0.f;
        RL_FEW_LOOP
        for (int i = 0; i < t.n_idx; ++i) {
          const float x = fminf(SMF(off, tc.idx_a[i]), 0.5f);
          const float d = x - mean;
          mean += div_count(d, i + 1);
          m2 += d * (x - mean);
        }
        r += div_count(m2, t.n_idx - 1);
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
      RL_FEW_LOOP
      for (int i = 0; i < t.n_idx; ++i) n += first_contact(sm, L, S, e, tc.idx_a[i]) ? 1 : 0;
      float r = ((float)n != t.p[0]) ? 1.f : 0.f;
      r *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_CONTACT_WITHOUT_CMD: {
      int n = 0;
      RL_FEW_LOOP
      for (int i = 0; i < t.n_idx; ++i) n += first_contact(sm, L, S, e, tc.idx_a[i]) ? 1 : 0;
      float r = (float)n;
      r *= (c.cmd_norm < 0.1f) ? 1.f : 0.f;
      return r * c.gate;
    }
    case RL_REW_FEET_STUMBLE: {
      bool any = false;   // t = 0 is the newest history sample = net_forces_w
      RL_TERM_LOOP
      for (int i = 0; i < t.n_idx; ++i) {
        const int b = tc.idx_c[i];
        const float fx = h[3 * b + 0], fy = h[3 * b + 1], fz = h[3 * b + 2];
        any = any || (sqrtf(fx * fx + fy * fy) > 4.f * fabsf(fz));
      }
      return (any ? 1.f : 0.f) * c.gate;
    }
    case RL_REW_FEET_SLIDE: {
      float s = 0.f;
      RL_FEW_LOOP
      for (int i = lo; i < t.n_idx && i < hi; ++i) {   // [lo, hi): this part's slice of the feet list
        const V3 vw = body_vec(sm, L, L.bvel, e, tc.idx_b[i]);
        const V3 vb = quat_apply_inverse(c.qw, c.q, V3{vw.x - c.vw.x, vw.y - c.vw.y, vw.z - c.vw.z});
        const float lat = sqrtf(vb.x * vb.x + vb.y * vb.y);
        s += lat * ((HIST_MAX_NORM(h, tc.idx_c[i]) > 1.0f) ? 1.f : 0.f);
      }
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT: {
      float s = 0.f;
      RL_FEW_LOOP
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 p = body_vec(sm, L, L.bpos, e, tc.idx_b[i]);
        const V3 v = body_vec(sm, L, L.bvel, e, tc.idx_b[i]);
        const float d = p.z - t.p[0];
        s += (d * d) * rl_tanhf(t.p[1] * sqrtf(v.x * v.x + v.y * v.y));
      }
      s *= (c.cmd_norm > 0.1f) ? 1.f : 0.f;
      return s * c.gate;
    }
    case RL_REW_FEET_HEIGHT_BODY: {
      float s = 0.f;
      RL_FEW_LOOP
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec(sm, L, L.bpos, e, tc.idx_b[i]);
        const V3 vw = body_vec(sm, L, L.bvel, e, tc.idx_b[i]);
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
      RL_TERM_LOOP
      for (int i = 0; i < t.n_idx; ++i) {
        const V3 pw = body_vec(sm, L, L.bpos, e, tc.idx_b[i]);
        const V3 pb = quat_apply_inverse(c.qw, c.q, V3{pw.x - c.pos.x, pw.y - c.pos.y, pw.z - c.pos.z});
        const float want = (t.p[0] / 2.f) * ((i % 2 == 0) ? 1.f : -1.f);
        const float d = want - pb.y;
        s += d * d;
      }
      return rl_expf(-s / t.p[1]) * c.gate;
    }
    case RL_REW_FEET_DISTANCE_XY_EXP: {
      float s = 0.f;
      RL_TERM_LOOP
      for (int i = 0; i < 4; ++i) {
        const V3 pw = body_vec(sm, L, L.bpos, e, tc.idx_b[i]);
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
      RL_TERM_LOOP
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
  c.g = quat_apply_inverse(c.qw, c.q, V3{0.f, 0.f, -1.f});
  c.vb = quat_apply_inverse(c.qw, c.q, c.vw);
  c.wb = quat_apply_inverse(c.qw, c.q, c.ww);
  c.gate = rl_div(clampf(-c.g.z, 0.f, 0.7f), 0.7f);   // through rl_div: a robot that is upside down has a zero numerator (special-operand path)
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
                                         const long long env, const EnvCtx& c, const int eplen_now,
                                         float* const out_sm = nullptr, const bool act_zero = false) {
  // out_sm: the record the row lives in when that is not `sm` (cluster kernels: another CTA's shared memory);
  // act_zero: ActionManager.reset [IL] has zeroed this env's stored action in this launch (not yet in the record)
  float* row = (out_sm ? out_sm : sm) + LOBS(g) + e * LOBSP(g);
  // noise-as-input mode (RlRandom.obs_uniforms: reproducibility hook for tests / replays, not the production
  // path): read straight from global memory
  const float* urow = a.rnd.obs_uniforms[g] ? a.rnd.obs_uniforms[g] + env * (g == 0 ? S.obs_dim0 : S.obs_dim1) : nullptr;
  const bool ext_u = (a.rnd.obs_uniforms[g] != nullptr);
  const bool noisy = t.has_noise && corrupt;
  RL_TERM_LOOP
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
        case RL_OBS_LAST_ACTION: v = act_zero ? 0.f : SMF(L.act, col); break;   // act_zero: this env's stored action was just reset
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

}  // namespace

#endif  // RL_MDP_TERMS_CUH_
