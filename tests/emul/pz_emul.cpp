// TEST INFRASTRUCTURE — the DEVICE code of pikazoo_b200/csrc (pz_state.cuh, pz_rng.cuh,
// pz_physics.cuh, and the samplers of pz_policy.cuh) compiled for the host with g++, one lane per "warp", so that the packing,
// the PCG64 restatement, the fast-forwarded trajectory simulations and the computer player can be
// fuzzed against the oracle in the `-m "not gpu"` suite (tests/test_device_code_on_host.py).
// It is never linked into the product: libpikazoo_b200.so is built by nvcc from the .cu files only,
// and nothing under pikazoo_b200/ can load this. The kernel glue (launch geometry, bulk-copy
// observation output, statistics atomics) is NOT covered here; the -m gpu tests cover it.
#define PZ_HOST_EMULATION 1
#include <cuda_runtime.h>  // vector types only; no CUDA call is made

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

// ---- one-lane warp shims -------------------------------------------------------------------------
struct Dim3Shim { unsigned x = 0, y = 0, z = 0; };
static Dim3Shim threadIdx_shim;
#define threadIdx threadIdx_shim
using std::max;
using std::min;
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __fns(unsigned mask, unsigned base, int offset) {
    int seen = 0;
    for (unsigned b = base; b < 32; b++)
        if (mask & (1u << b))
            if (++seen == offset) return (int)b;
    return -1;
}
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
template <typename T>
static inline T __ldg(const T *p) { return *p; }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int __ffs(int v) { return __builtin_ffs(v); }

static inline float __uint_as_float(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t v; std::memcpy(&v, &f, 4); return v; }

#include "../../pikazoo_b200/csrc/pz_physics.cuh"
#include "../../pikazoo_b200/csrc/pz_policy.cuh"

using namespace pz;

namespace {

struct Batch {
    int64_t n;
    std::vector<int32_t> state;  // the packed SoA buffer, exactly as on the device
};

template <int AI_MASK>
int step_one(Env &e, DrawCtx &d, const StepCfg &c, Input in1, Input in2, int *scratch) {
    return step_frame_inputs<AI_MASK>(1u, e, d, c, in1, in2, scratch);
}

}  // namespace

extern "C" {

int emul_state_words() { return 17; }

// decode_input (the kernels' one-step decode) against get_input(decode_keys) (the table form the policy rollout
// pre-decodes from): out[0..4] = xdir, ydir, power, keyprev after, bad of the former, out[5..9] of the latter
void emul_decode_both(int agent, int simplify, int action, int keyprev, int32_t *out) {
    for (int form = 0; form < 2; form++) {
        Player p = {};
        p.keyprev = keyprev;
        bool bad = false;
        Input in;
        if (form == 0) {
            in = agent == 0 ? (simplify ? decode_input<0, true>(action, p, bad) : decode_input<0, false>(action, p, bad))
                            : (simplify ? decode_input<1, true>(action, p, bad) : decode_input<1, false>(action, p, bad));
        } else {
            const uint32_t keys = agent == 0 ? (simplify ? decode_keys<0, true>(action, bad) : decode_keys<0, false>(action, bad))
                                             : (simplify ? decode_keys<1, true>(action, bad) : decode_keys<1, false>(action, bad));
            in = get_input(p, keys);
        }
        int32_t *o = out + 5 * form;
        o[0] = in.xdir, o[1] = in.ydir, o[2] = in.power, o[3] = p.keyprev, o[4] = bad;
    }
}

// unpacked int32[n][53] -> packed SoA (same conversion as pz_import_state)
void emul_import(int32_t *packed, int64_t n, const int32_t *unpacked) {
    StatePtrs s = state_ptrs(packed, n);
    for (int64_t i = 0; i < n; i++) {
        const int32_t *o = unpacked + i * 53;
        Env e;
        env_from_unpacked(e, o);
        store_env(e, s, i);
        s.g2[i] = make_int4(o[42], o[43], o[44], o[45]);
        s.g3[i] = make_int4(o[46], o[47], o[48], o[49]);
        s.u[i] = (uint32_t)o[51];
    }
}

void emul_export(const int32_t *packed, int64_t n, int32_t *unpacked) {
    StatePtrs s = state_ptrs(const_cast<int32_t *>(packed), n);
    for (int64_t i = 0; i < n; i++) {
        int32_t *o = unpacked + i * 53;
        Env e;
        load_env(e, s, i);
        env_to_unpacked(e, o);
        const int4 st = s.g2[i], ic = s.g3[i];
        o[42] = st.x, o[43] = st.y, o[44] = st.z, o[45] = st.w;
        o[46] = ic.x, o[47] = ic.y, o[48] = ic.z, o[49] = ic.w;
        o[51] = (int32_t)s.u[i];
    }
}

void emul_seed(int32_t *packed, int64_t n, uint64_t base_seed) {
    StatePtrs s = state_ptrs(packed, n);
    for (int64_t i = 0; i < n; i++) {
        Env e;
        fresh_env(e);
        Rng r;
        pcg64_seed(base_seed + (uint64_t)i, r);
        store_env(e, s, i);
        rng_store(r, s, i);
        s.g3[i] = make_int4((int)(uint32_t)r.inc_lo, (int)(uint32_t)(r.inc_lo >> 32), (int)(uint32_t)r.inc_hi,
                            (int)(uint32_t)(r.inc_hi >> 32));
    }
}

// One batched call with the product's NEXT-STEP auto-reset semantics (mirrors pz_step_kernel's body):
// base reward of player_1 in {-1,0,1} per env, obs int32[n][2][35], done u8[n]. Every env goes
// through a load_env/store_env round trip, so the packed layout is exercised on every frame.
void emul_step(int32_t *packed, int64_t n, int winning_score, int serve, int ai_mask, int simplify, int autoreset,
               const int32_t *actions, int32_t *obs, int32_t *base_reward, uint8_t *done, int do_reset_all) {
    StatePtrs s = state_ptrs(packed, n);
    StepCfg c;
    c.winning_score = winning_score;
    c.serve = serve;
    c.tab_land = nullptr;
    c.tab_power = nullptr;
    static int scratch[32 * 70];
    for (int64_t i = 0; i < n; i++) {
        DrawCtx d;
        d.s = s;
        d.idx = i;
        d.r.loaded = false;
        d.r.dirty = false;
        Env e;
        load_env(e, s, i);
        int base = 0;
        bool stepped = false;
        if (do_reset_all) {
            reset_env(e, d, c);
        } else if (!e.game_ended) {
            const int a1 = actions ? actions[2 * i] : 0, a2 = actions ? actions[2 * i + 1] : 0;
            bool b1, b2;
            Input k1, k2;  // the kernels' decode: action -> Input in one go
            if (simplify) {
                k1 = decode_input<0, true>(a1, e.p[0], b1);
                k2 = decode_input<1, true>(a2, e.p[1], b2);
            } else {
                k1 = decode_input<0, false>(a1, e.p[0], b1);
                k2 = decode_input<1, false>(a2, e.p[1], b2);
            }
            if (ai_mask != 0) rng_load(d.r, s, i);
            switch (ai_mask) {
                case 0: base = step_one<0>(e, d, c, k1, k2, scratch); break;
                case 1: base = step_one<1>(e, d, c, k1, k2, scratch); break;
                case 2: base = step_one<2>(e, d, c, k1, k2, scratch); break;
                default: base = step_one<3>(e, d, c, k1, k2, scratch); break;
            }
            stepped = true;
        } else if (autoreset) {
            reset_env(e, d, c);
        }
        if (do_reset_all || stepped || autoreset) {
            store_env(e, s, i);
            if (d.r.dirty) rng_store(d.r, s, i);
        }
        if (obs) {
            int u[35];
            obs_values(e, u);
            for (int k = 0; k < 70; k++) obs[i * 70 + k] = u[obs_src(k)];
        }
        if (base_reward) base_reward[i] = base;
        if (done) done[i] = (uint8_t)((stepped && e.game_ended) || (!stepped && !do_reset_all && !autoreset));
    }
}

// player_move with the sprite animation as a table (what the K-frame kernels read from shared memory after anim_fill,
// and most per-step kernels from the compile-time g_anim_table) against the arithmetic form, over every animation
// state x a grid of positions, velocities and inputs. Returns the number of combinations checked; *mismatches counts
// the ones where any player field differs (between the three forms).
int64_t emul_player_move_forms(int64_t *mismatches) {
    static uint32_t filled[kAnimLutEntries];
    anim_fill(filled, 0, 1);
    int64_t checked = 0, bad = 0;
    const int ys[] = {kPlayerGroundY, kPlayerGroundY - 1, kPlayerGroundY - 16, 108, kPlayerGroundY + 3};
    const int yvs[] = {-16, -5, -1, 0, 1, 7, 16};
    for (int state = 0; state < 5; state++)
        for (int frame = 0; frame < 8; frame++)
            for (int delay = 0; delay < 8; delay++)
                for (int arm = -1; arm <= 1; arm += 2)
                    for (int y : ys)
                        for (int yv : yvs)
                            for (int lying = -2; lying <= 3; lying++)
                                for (int dive = -1; dive <= 1; dive++)
                                    for (int inp = 0; inp < 18; inp++)
                                        for (int x : {32, 100, 184, 248, 400}) {
                                            Player p = {};
                                            p.x = x, p.y = y, p.yv = yv, p.state = state, p.frame = frame, p.delay = delay;
                                            p.arm = arm, p.dive = dive, p.lying = lying, p.bold = 3, p.standby = 1;
                                            Input in;
                                            in.xdir = inp % 3 - 1, in.ydir = (inp / 3) % 3 - 1, in.power = inp / 9;
                                            Player a = p, b = p, c = p, a1 = p, b1 = p;
                                            player_move<0, false>(a, in);
                                            player_move<0, true>(b, in, filled);
                                            player_move<0, true>(c, in, g_anim_table.v);
                                            player_move<1, false>(a1, in);
                                            player_move<1, true>(b1, in, filled);
                                            checked++;
                                            if (std::memcmp(&a, &b, sizeof a) || std::memcmp(&a, &c, sizeof a) ||
                                                std::memcmp(&a1, &b1, sizeof a1))
                                                bad++;
                                        }
    *mismatches = bad;
    return checked;
}

// The two trajectory simulations alone (fast-forwarded device form), for exhaustive comparison
// with the reference's plain loops. Returns landing x; *by_ground = ended on the ground.
int emul_simulate(int x, int y, int xv, int yv, int power, int *by_ground) {
    bool g;
    const int lx = power ? simulate_landing_x<true>(1u, x, y, xv, yv, true, g)
                         : simulate_landing_x<false>(1u, x, y, xv, yv, true, g);
    if (by_ground) *by_ground = g ? 1 : 0;
    return lx;
}

void emul_simulate_many(int64_t n, const int32_t *xyv, int power, int32_t *out) {
    for (int64_t i = 0; i < n; i++) {
        bool g;
        const int32_t *q = xyv + 4 * i;
        out[i] = power ? simulate_landing_x<true>(1u, q[0], q[1], q[2], q[3], true, g)
                       : simulate_landing_x<false>(1u, q[0], q[1], q[2], q[3], true, g);
    }
}

// NormalizeObservation as the kernels compute it: u int32[n][35] -> float32 / float64 [n][35]
void emul_normalize(int64_t n, const int32_t *u, float *f32, double *f64, int normalize) {
    for (int64_t i = 0; i < n; i++) {
        int v[35];
        for (int k = 0; k < 35; k++) v[k] = u[i * 35 + k];
        float a[35];
        double b[35];
        obs_floats(v, a, normalize != 0);
        obs_floats(v, b, normalize != 0);
        for (int k = 0; k < 35; k++) f32[i * 35 + k] = a[k], f64[i * 35 + k] = b[k];
    }
}

int emul_synth_action(uint64_t seed, uint64_t env, uint64_t frame, int agent, uint32_t n_actions) {
    return synth_action(seed, env, frame, agent, n_actions);
}

// The policy kernels' samplers (pz_policy.cuh) on host logits float32 [n][2][n_actions]:
//   emul_policy_inverse_cdf: what the tcgen05 kernel's threads compute (NA = 18 instantiation when n_actions == 18,
//                            the generic 24-candidate one otherwise), actions int32 [n][2]
//   emul_policy_gumbel_keys: the mma.sync kernel's keys logit + noise (before packing), float32 [n][2][n_actions]
void emul_policy_inverse_cdf(int64_t n, int n_actions, const float *logits, uint64_t seed, uint64_t step,
                             uint64_t first_env, int generic, int32_t *actions) {
    for (int64_t e = 0; e < n; e++) {
        const uint32_t nbase = pzp::noise_base(seed, step, first_env + (uint64_t)e);
        for (int a = 0; a < 2; a++) {
            const float *l = logits + (e * 2 + a) * n_actions;
            const uint32_t agent_base = nbase + (uint32_t)(32 * a) * 0x9E3779B9u;
            if (n_actions == 18 && !generic) {
                float v[18];
                for (int j = 0; j < 18; j++) v[j] = l[j];
                actions[e * 2 + a] = pzp::sample_inverse_cdf<18>(v, 18, agent_base);
            } else {
                float v[PZ_POLICY_MAX_ACTIONS];
                for (int j = 0; j < PZ_POLICY_MAX_ACTIONS; j++) v[j] = j < n_actions ? l[j] : 0.0f;
                actions[e * 2 + a] = pzp::sample_inverse_cdf<PZ_POLICY_MAX_ACTIONS>(v, n_actions, agent_base);
            }
        }
    }
}

void emul_policy_gumbel_keys(int64_t n, int n_actions, const float *logits, uint64_t seed, uint64_t step,
                             uint64_t first_env, float *keys) {
    for (int64_t e = 0; e < n; e++) {
        const uint32_t nbase = pzp::noise_base(seed, step, first_env + (uint64_t)e);
        for (int a = 0; a < 2; a++)
            for (int j = 0; j < n_actions; j++) {
                const int64_t i = (e * 2 + a) * n_actions + j;
                keys[i] = pzp::gumbel_key(logits[i], nbase, a, j);
            }
    }
}

}  // extern "C"
