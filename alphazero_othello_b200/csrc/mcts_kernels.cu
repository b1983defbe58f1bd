// Batched MCTS self-play engine for B200 (sm_100a).
//
// One group of LANES threads (8 by default; 16 / 32 selectable) owns one game
// slot.  The tree of a slot lives in two flat arenas in HBM (32-byte node
// records + 16-byte boards); children of a node are contiguous and in
// ascending action order, so PUCT selection is one 2x128-bit load per lane and
// a two-instruction warp-reduce arg-max.  k_mcts_step (hot: expand, backup,
// descend) hands slots whose move is due to k_mcts_move (policy target, move
// sampling, re-rooting = breadth-first compaction of the kept subtree into the
// other arena, game hand-off), both launched by one oth_mcts_step call.
//
// Semantics restated from the reference (num_threads = 1):
//   Node / PUCT / expand / backup    MCTS_model.py:46-169
//   MCTS.policy_improve_step etc.    MCTS_model.py:172-395
//   one_self_play, lambda-returns    self_play_worker.py:8-88
// including numpy>=2 promotion: float32 PUCT for ordinary nodes, float64 at a
// Dirichlet-noised root, float64 value sums, numpy's 8-lane pairwise sums.
// Compiled with --fmad=false; the rounding-critical steps also use the
// explicit *_rn intrinsics.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include "../../include/othello_b200.h"
#include "bitboard.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace cg = cooperative_groups;
using namespace oth;

namespace {

constexpr int kBlock = 128;
#ifndef OTH_STEP_BLOCKS_PER_SM
#define OTH_STEP_BLOCKS_PER_SM 8
#endif
constexpr int kStepBlocksPerSM = OTH_STEP_BLOCKS_PER_SM;  // resident step-kernel blocks per SM the register budget is sized for
constexpr int kMoveSet = 8;  // slots per warp in k_mcts_move
constexpr uint64_t kSaltNoise = 0x6e6f697365ULL;  // Philox purposes
constexpr uint32_t kPurposeMove = 1, kPurposeTie = 2;
constexpr uint64_t kSaltRollout = 0x726f6c6c6f7574ULL;

struct __align__(32) Node {
    double W;             // value_sum (MCTS_model.py:84), float64
    float prior;          // float32 prior (float64 copy in root_prior64 at a noised root)
    int32_t N;            // visit_count
    int32_t first_child;  // -1: leaf
    uint32_t meta;        // nchild | action << 8 | flags << 16 | (terminal_value & 0xff) << 24
    u64 moves;            // legal set of the side to move at this node (Node.valid_mask :96)
};
static_assert(sizeof(Node) == 32, "node record must be one 32-byte sector");

constexpr uint32_t kFlagTerminal = 1u;

__device__ __forceinline__ int meta_nchild(uint32_t m) { return (int)(m & 0xffu); }
__device__ __forceinline__ int meta_action(uint32_t m) { return (int)((m >> 8) & 0xffu); }
__device__ __forceinline__ uint32_t meta_flags(uint32_t m) { return (m >> 16) & 0xffu; }
__device__ __forceinline__ int meta_tvalue(uint32_t m) { return (int)(int8_t)(m >> 24); }
__device__ __forceinline__ uint32_t make_meta(int nchild, int action, uint32_t flags, int tv)
{
    return (uint32_t)nchild | ((uint32_t)action << 8) | (flags << 16) | (((uint32_t)tv & 0xffu) << 24);
}

// Two 128-bit accesses per record, always served by L2 (ld.global.cg): visit counts and value
// sums are updated by fire-and-forget L2 reductions (backup), so no record line may sit in L1.
__device__ __forceinline__ uint4 ldcg4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ Node load_node(const Node* p)
{
    const uint4 a = ldcg4(p);
    const uint4 b = ldcg4(reinterpret_cast<const uint4*>(p) + 1);
    Node n;
    n.W = __hiloint2double((int)a.y, (int)a.x);
    n.prior = __uint_as_float(a.z);
    n.N = (int)a.w;
    n.first_child = (int)b.x;
    n.meta = b.y;
    n.moves = ((u64)b.w << 32) | b.z;
    return n;
}

__device__ __forceinline__ void store_node(Node* p, const Node& n)
{
    uint4 a, b;
    a.x = (unsigned)__double2loint(n.W);
    a.y = (unsigned)__double2hiint(n.W);
    a.z = __float_as_uint(n.prior);
    a.w = (unsigned)n.N;
    b.x = (unsigned)n.first_child;
    b.y = n.meta;
    b.z = (unsigned)n.moves;
    b.w = (unsigned)(n.moves >> 32);
    *reinterpret_cast<uint4*>(p) = a;
    *(reinterpret_cast<uint4*>(p) + 1) = b;
}

struct Params {
    oth_mcts_config cfg;
    Node* nodes;
    ulonglong2* boards;
    oth_mcts_ctl* ctl;
    int* path;
    double* root_prior64;
    double* noise;
    double* u_move;
    double* u_tie;
    ulonglong2* traj_board;
    float* traj_pi;
    double* traj_rootv;
    int* traj_meta;
    ulonglong2* out_board;
    float* out_pi;
    double* out_value;
    long long* out_meta;
    long long* out_games;
    unsigned long long* counters;
    unsigned* slot_counters;  // [slot][16] cumulative event counters, summed by k_mcts_poll
    uint4* hot;               // [slot] 256-byte SlotHot records
    uint8_t* move_flags;      // [slot] set by the step kernel when a slot's move is due (k_mcts_move clears it)
    int* move_list;           // [0] slots due, [1] block tickets, [4..] due slots (cfg.move_launch = 1)
    int hot_path;             // path entries kept in the hot record (cfg.hot_path, default kHotPath)
    const float* priors;
    const float* values;
    // fused network tail (oth_mcts_step_fused): raw logits / value pre-activations, row strides in elements
    const void* logits;
    const void* vpre;
    int logits_stride, vpre_stride, raw_bf16;
    const int* eval_map;  // optional (evaluation de-duplication): row of logits / vpre that holds slot s's evaluation, -1 = none this launch
    float* priors_out;  // optional: the softmax / tanh the kernel applied, for record & replay
    float* values_out;
    float* nn_input;
    float c_puct_f32;
};

// per-group scratch in shared memory
enum { CNT_LOCAL = 16 };

// Everything a slot needs at the start of a launch, in one 256-byte record fetched together
// with the control block: the pending leaf (board, legal set, meta word) so expansion can start
// without a second round trip, a mirror of the root's header so the descent starts without
// reading the root record, and the first 52 entries of the search path.
constexpr int kHotPath = 52;
struct __align__(16) SlotHot {
    u64 leaf_own, leaf_opp, leaf_moves;
    uint32_t leaf_meta;
    int32_t root_N;  // == visit_count of the root record
    int32_t root_fc;  // == first_child of the root record (-1: leaf)
    uint32_t root_meta;
    int32_t pad[2];
    int32_t path[kHotPath];
};
static_assert(sizeof(SlotHot) == 256 && offsetof(SlotHot, path) == 48, "SlotHot layout");

struct __align__(16) Scratch {
    double pri64[OTH_NUM_ACTIONS + 1];
    float pri[OTH_NUM_ACTIONS + 3];
    SlotHot hot;               // hot.path continues into path_tail: 52 + 76 = 128 entries
    int path_tail[128 - kHotPath];
    unsigned cnt[CNT_LOCAL];  // the slot's cumulative event counters while it is being worked on
};
static_assert(offsetof(Scratch, path_tail) == offsetof(Scratch, hot) + 256, "path must be contiguous");

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// sm_100a warp-reduce units: float max and integer min over the lanes named by `mask`
__device__ __forceinline__ float redux_max_f32(float v, unsigned mask)
{
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(m) : "f"(v), "r"(mask));
    return m;
}

template <int LANES>
struct Ctx {
    cg::thread_block_tile<LANES> tile;
    const Params& P;
    Scratch& S;
    int lane;
    unsigned gmask;  // lanes of this group within the warp
    int slot;
    Node* N;        // current arena
    ulonglong2* B;
    oth_mcts_ctl c;

    __device__ Ctx(cg::thread_block_tile<LANES> t, const Params& p, Scratch& s) : tile(t), P(p), S(s), lane(t.thread_rank()), slot(0)
    {
        gmask = LANES == 32 ? 0xffffffffu : (((1u << (LANES & 31)) - 1u) << ((threadIdx.x & 31) & ~(LANES - 1)));
    }

    // per-slot event counters live in HBM next to the control block: no atomics, no block barrier
    __device__ __forceinline__ void load_counters()
    {
        for (int i = lane; i < CNT_LOCAL; i += LANES) S.cnt[i] = P.slot_counters[(size_t)slot * CNT_LOCAL + i];
    }
    __device__ __forceinline__ void store_counters()
    {
        gsync();
        for (int i = lane; i < CNT_LOCAL; i += LANES) P.slot_counters[(size_t)slot * CNT_LOCAL + i] = S.cnt[i];
    }

    // group collectives on the precomputed lane mask (cheaper than the cooperative-groups tile wrappers)
    __device__ __forceinline__ void gsync() const { __syncwarp(gmask); }
    template <typename T>
    __device__ __forceinline__ T gshfl(T v, int src) const { return __shfl_sync(gmask, v, src, LANES); }
    template <typename T>
    __device__ __forceinline__ T gshfl_xor(T v, int m) const { return __shfl_xor_sync(gmask, v, m, LANES); }
    template <typename T>
    __device__ __forceinline__ T gshfl_up(T v, int d) const { return __shfl_up_sync(gmask, v, d, LANES); }
    __device__ __forceinline__ unsigned gballot(bool p) const
    {
        return LANES == 32 ? __ballot_sync(gmask, p) : ((__ballot_sync(gmask, p) & gmask) >> ((threadIdx.x & 31) & ~(LANES - 1)));
    }

    __device__ __forceinline__ void count(int which, unsigned n = 1)
    {
        if (lane == 0) S.cnt[which] += n;
    }
    __device__ __forceinline__ void count_max(int which, unsigned v)
    {
        if (lane == 0 && v > S.cnt[which]) S.cnt[which] = v;
    }

    // first 80 bytes (header + 8 path entries) are requested with the control block; the rest of
    // the path only if it is that deep
    __device__ __forceinline__ void load_hot_head()
    {
        const uint4* g = P.hot + (size_t)slot * 16;
        uint4* d = reinterpret_cast<uint4*>(&S.hot);
        for (int i = lane; i < 5; i += LANES) d[i] = g[i];
    }
    __device__ __forceinline__ void load_path_rest(int path_len)
    {
        const int hp = P.hot_path;  // entries [0, hp) travel in the hot record, [hp, path_len) in OTH_BUF_PATH
        if (path_len > 8) {
            const uint4* g = P.hot + (size_t)slot * 16;
            uint4* d = reinterpret_cast<uint4*>(&S.hot);
            const int n16 = 3 + (min(path_len, hp) + 3) / 4;
            for (int i = 5 + lane; i < n16; i += LANES) d[i] = g[i];
        }
        if (path_len > hp) {
            const int* gp = P.path + (size_t)slot * P.cfg.path_cap;
            for (int k = hp + lane; k < path_len; k += LANES) S.hot.path[k] = gp[k];
        }
    }
    __device__ __forceinline__ void store_hot(int path_len)
    {
        gsync();
        const int hp = P.hot_path;
        uint4* g = P.hot + (size_t)slot * 16;
        const uint4* d = reinterpret_cast<const uint4*>(&S.hot);
        const int n16 = 3 + (min(path_len, hp) + 3) / 4;
        for (int i = lane; i < n16; i += LANES) g[i] = d[i];
        if (path_len > hp) {
            int* gp = P.path + (size_t)slot * P.cfg.path_cap;
            for (int k = hp + lane; k < path_len; k += LANES) gp[k] = S.hot.path[k];
        }
    }

    __device__ __forceinline__ void bind_arena()
    {
        const size_t base = ((size_t)slot * 2 + (size_t)c.arena) * (size_t)P.cfg.node_cap;
        N = P.nodes + base;
        B = P.boards + base;
    }

    __device__ __forceinline__ void fail(int bit)
    {
        c.error |= bit;
        c.phase = OTH_PH_ERROR;
    }

    // -DOTH_DEBUG build: every arena index is checked against the slot's bump pointer / capacity before use.
    // A violation is recorded (OTH_ERR_DEBUG, source line in ctl.reserved) and the index clamped to the root,
    // so the launch finishes and the host can report it instead of taking an illegal-address fault.
#ifdef OTH_DEBUG
    __device__ __noinline__ int dbg_idx(int i, int lim, int line)
    {
        if ((unsigned)i >= (unsigned)lim) {
            c.error |= OTH_ERR_DEBUG;
            c.reserved = ((long long)line << 32) | (unsigned)i;
            return 0;
        }
        return i;
    }
#define OTH_IDX(i, lim) dbg_idx((i), (lim), __LINE__)
#else
#define OTH_IDX(i, lim) (i)
#endif

    // numpy pairwise add.reduce over 65 float32 in S.pri (MCTS_model.py:347, :259)
    __device__ __forceinline__ float np_sum65_f32()
    {
        float r = 0.0f;
        if (lane < 8) {
            r = S.pri[lane];
#pragma unroll
            for (int i = 1; i < 8; i++) r = __fadd_rn(r, S.pri[8 * i + lane]);
        }
        r = __fadd_rn(r, gshfl_xor(r, 1));
        r = __fadd_rn(r, gshfl_xor(r, 2));
        r = __fadd_rn(r, gshfl_xor(r, 4));
        r = gshfl(r, 0);
        return __fadd_rn(r, S.pri[64]);
    }

    __device__ __forceinline__ double np_sum65_f64()
    {
        double r = 0.0;
        if (lane < 8) {
            r = S.pri64[lane];
#pragma unroll
            for (int i = 1; i < 8; i++) r = __dadd_rn(r, S.pri64[8 * i + lane]);
        }
        r = __dadd_rn(r, gshfl_xor(r, 1));
        r = __dadd_rn(r, gshfl_xor(r, 2));
        r = __dadd_rn(r, gshfl_xor(r, 4));
        r = gshfl(r, 0);
        return __dadd_rn(r, S.pri64[64]);
    }

    // canonical plane of the side to move: +1 own, -1 opp (Models.py:16)
    __device__ __forceinline__ void write_nn_input(u64 own, u64 opp)
    {
        float* dst = P.nn_input + (size_t)slot * 64;
        for (int e = lane; e < 64; e += LANES) dst[e] = ((own >> e) & 1) ? 1.0f : (((opp >> e) & 1) ? -1.0f : 0.0f);
    }

    // ------------------------------------------------------------ stubs --
    __device__ __forceinline__ u64 mix64(u64 x)
    {
        x ^= x >> 33;
        x *= 0xff51afd7ed558ccdULL;
        x ^= x >> 33;
        x *= 0xc4ceb9fe1a85ec53ULL;
        x ^= x >> 33;
        return x;
    }

    // policy=None evaluation (MCTS_model.py:332-335): uniform priors np.ones(65) and the outcome of
    // one uniform-random playout from the leaf, seen from the leaf's side to move (_rollout
    // :276-303).  Draws come from Philox keyed (seed, game id, ply/root flag, simulation index,
    // 4 plies per block); move choice = floor(u32 * n_legal / 2^32)-th legal action, ascending.
    __device__ double eval_rollout(u64 own, u64 opp, bool root_init)
    {
        for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) S.pri[a] = 1.0f;
        int result = 0;
        if (lane == 0) {
            const uint32_t tag = ((uint32_t)c.ply << 1) | (root_init ? 1u : 0u);
            const uint32_t ctr = (uint32_t)c.sims_done;
            u64 o = own, p = opp;
            int side = 0;  // 0: the leaf's side to move owns `o`
            u64 m = legal_moves(o, p);
            Philox4 r = {0, 0, 0, 0};
            for (int nd = 0; nd < 128;) {  // nd = random draws so far (forced passes draw nothing)
                if (m == 0) {
                    const u64 m2 = legal_moves(p, o);
                    if (m2 == 0) break;  // neither side can move: terminal
                    const u64 t = o;
                    o = p;
                    p = t;
                    side ^= 1;
                    m = m2;
                    continue;
                }
                if ((nd & 3) == 0) r = philox4x32_10(P.cfg.seed ^ kSaltRollout, (uint64_t)c.game_id, (uint32_t)ctr, tag | ((uint32_t)(nd >> 2) << 16));
                const int w = nd & 3;
                const uint32_t rnd = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
                nd++;
                const int sq = nth_set_bit(m, (int)__umulhi(rnd, (uint32_t)__popcll(m)));
                const u64 f = flips(o, p, 1ULL << sq);
                const Board b = apply_move(o, p, sq, f);
                o = b.own;
                p = b.opp;
                side ^= 1;
                m = legal_moves(o, p);
            }
            const int d = side == 0 ? __popcll(o) - __popcll(p) : __popcll(p) - __popcll(o);
            result = (d > 0) - (d < 0);
        }
        result = gshfl(result, 0);
        gsync();
        return (double)result;
    }

    // Device twins of oracle/othello_oracle.c orc_stub_a/b/h: raw priors to S.pri, value returned.
    __device__ double eval_stub(u64 own, u64 opp)
    {
        double value = 0.0;
        const int kind = P.cfg.eval_kind;
        if (kind == OTH_EVAL_STUB_A) {
            const float v = (float)(1.0 / 65.0);
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) S.pri[a] = v;
        } else if (kind == OTH_EVAL_STUB_B) {
            long long h = 0;
            for (int e = lane; e < 64; e += LANES) h += ((own >> e) & 1) ? (e + 1) : (((opp >> e) & 1) ? -(e + 1) : 0);
            for (int o = LANES / 2; o; o >>= 1) h += gshfl_xor(h, o);
            float sum = 0.0f;  // small integers: exact in any order
            for (int a = 0; a < OTH_NUM_ACTIONS; a++) {
                long long t = (7LL * a + h) % 11;
                if (t < 0) t += 11;
                sum += (float)(t + 1);
            }
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) {
                long long t = (7LL * a + h) % 11;
                if (t < 0) t += 11;
                S.pri[a] = __fdiv_rn((float)(t + 1), sum);
            }
            long long t = h % 17;
            if (t < 0) t += 17;
            value = (double)(t - 8) / 16.0;
        } else {  // OTH_EVAL_STUB_H: hashed in the reference's bit numbering (bit = 63 - idx)
            const u64 h = mix64(__brevll(own) ^ mix64(__brevll(opp) ^ P.cfg.stub_salt));
            int part = 0;
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) part += 1 + (int)(mix64(h + (u64)a) % 251ULL);
            for (int o = LANES / 2; o; o >>= 1) part += gshfl_xor(part, o);
            const float sum = (float)part;
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES)
                S.pri[a] = __fdiv_rn((float)(1 + (int)(mix64(h + (u64)a) % 251ULL)), sum);
            const int v = (int)(mix64(h ^ 0x9e3779b97f4a7c15ULL) % 2001ULL) - 1000;
            value = (double)__fdiv_rn((float)v, 1000.0f);
        }
        gsync();
        return value;
    }

    // ------------------------------------------------- Dirichlet noise --
    // np.random.dirichlet([alpha]*65) (MCTS_model.py:341) drawn from Philox:
    // Marsaglia-Tsang gammas, normalised.  The draw is written to P.noise so
    // a checker can replay it; with inject_random it is read from there.
    __device__ void make_noise()
    {
        double* nz = P.noise + (size_t)slot * OTH_NUM_ACTIONS;
        if (!P.cfg.inject_random) {
            const double alpha = P.cfg.dirichlet_alpha;
            const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
            const double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
            for (int e = lane; e < OTH_NUM_ACTIONS; e += LANES) {
                double g = d;
                for (uint32_t k = 0; k < 256; k++) {
                    // ply 0 (game start) keeps the round-1 stream; a later fresh root (manual mode) is keyed by its ply
                    const Philox4 r = philox4x32_10(P.cfg.seed ^ kSaltNoise, (uint64_t)c.game_id, (uint32_t)e, k | ((uint32_t)c.ply << 16));
                    const double u1 = ((double)r.x + 1.0) * (1.0 / 4294967296.0);
                    const double u2 = (double)r.y * (1.0 / 4294967296.0);
                    const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
                    double v = 1.0 + cc * x;
                    if (v <= 0.0) continue;
                    v = v * v * v;
                    const double u = ((double)r.z + 0.5) * (1.0 / 4294967296.0);
                    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
                        g = d * v;
                        if (alpha < 1.0) g *= pow(((double)r.w + 0.5) * (1.0 / 4294967296.0), 1.0 / alpha);
                        break;
                    }
                }
                S.pri64[e] = g;
            }
            gsync();
            double s = 0.0;
            for (int e = 0; e < OTH_NUM_ACTIONS; e++) s += S.pri64[e];
            gsync();
            for (int e = lane; e < OTH_NUM_ACTIONS; e += LANES) nz[e] = S.pri64[e] / s;
            gsync();
        }
    }

    // -------------------------------------------------------- expansion --
    // MCTS._expand_and_evaluate (MCTS_model.py:325-360) given raw float32
    // priors in S.pri: noise mix (root only), mask, pairwise-sum normalise,
    // create every child in ascending action order (Node.expand :146-158,
    // Node.__init__ :68-108 -- child boards, legal sets and terminal values
    // are computed here, one child per lane).
    __device__ bool expand(int leaf, bool root_init, const Node& lf, const ulonglong2 lb, bool premasked)
    {
        const u64 own = lb.x, opp = lb.y;
        const u64 M = lf.moves;
        const bool is_pass = (M == 0);
        const int nchild = is_pass ? 1 : __popcll(M);
        const int fc = c.top;
        if (fc + nchild > P.cfg.node_cap) {
            fail(OTH_ERR_NODE_OVERFLOW);
            return false;
        }
        const bool f64 = root_init && P.cfg.dirichlet_epsilon > 0.0;
        float sum32 = 0.0f;
        if (f64) {
            const double* nz = P.noise + (size_t)slot * OTH_NUM_ACTIONS;  // drawn when the tree was created (init_tree) or injected
            const double eps = P.cfg.dirichlet_epsilon;
            const float om = (float)(1.0 - eps);  // (1-eps)*priors stays float32 (NEP 50)
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) {
                const float t = __fmul_rn(om, S.pri[a]);
                double v = __dadd_rn((double)t, __dmul_rn(eps, nz[a]));
                const bool valid = a < 64 ? ((M >> a) & 1) : is_pass;
                S.pri64[a] = valid ? v : __dmul_rn(v, 0.0);
            }
            gsync();
            const double s = np_sum65_f64();
            gsync();
            if (!isfinite(s)) {
                fail(OTH_ERR_NONFINITE);
                return false;
            }
            if (s > 1e-12)
                for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) S.pri64[a] = __ddiv_rn(S.pri64[a], s);
        } else {
            if (!premasked) {
                for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) {
                    const bool valid = a < 64 ? ((M >> a) & 1) : is_pass;
                    if (!valid) S.pri[a] = __fmul_rn(S.pri[a], 0.0f);
                }
                gsync();
            }
            sum32 = np_sum65_f32();  // the division (priors /= sum, :348-349) is applied per child below
            if (!isfinite(sum32)) {  // a NaN / infinite prior: stop this slot, never let it into the tree
                fail(OTH_ERR_NONFINITE);
                return false;
            }
        }
        gsync();
        for (int i = lane; i < nchild; i += LANES) {
            const int a = is_pass ? OTH_PASS : nth_set_bit(M, i);
            const u64 f = is_pass ? 0ULL : flips(own, opp, 1ULL << a);
            const Board cb = apply_move(own, opp, a, f);
            Node ch;
            ch.W = 0.0;
            ch.N = 0;
            ch.first_child = -1;
            ch.moves = legal_moves(cb.own, cb.opp);
            int tv;
            const int st = position_status(cb.own, cb.opp, ch.moves, &tv);
            ch.meta = make_meta(0, a, st == 2 ? kFlagTerminal : 0u, tv);
            if (f64) {
                P.root_prior64[(size_t)slot * OTH_MAX_CHILDREN + i] = S.pri64[a];
                ch.prior = (float)S.pri64[a];
            } else {
                ch.prior = sum32 > (float)1e-12 ? __fdiv_rn(S.pri[a], sum32) : S.pri[a];
            }
            store_node(N + OTH_IDX(fc + i, P.cfg.node_cap), ch);
            B[fc + i] = make_ulonglong2(cb.own, cb.opp);
        }
        if (lane == 0) {
            Node* p = N + OTH_IDX(leaf, c.top);
            p->first_child = fc;
            p->meta = (lf.meta & ~0xffu) | (uint32_t)nchild;
        }
        c.top = fc + nchild;
        if (f64) c.flags |= 2;
        if (leaf == c.root && lane == 0) {  // keep the root mirror exact
            S.hot.root_fc = fc;
            S.hot.root_meta = (lf.meta & ~0xffu) | (uint32_t)nchild;
        }
        count(OTH_CNT_NODES, nchild);
        gsync();
        return true;
    }

    // Node.backpropagate (MCTS_model.py:160-169) along S.hot.path[0..depth).
    __device__ __forceinline__ void backup(int depth, double value)
    {
        gsync();
        // one float64 add and one integer add per path node, performed by the L2 (RED, no return
        // value): nothing is loaded, the group does not wait.  A single IEEE add per node per
        // simulation, in simulation order -> same sums as the reference's sequential loop.
        for (int d = lane; d < depth; d += LANES) {
            Node* p = N + OTH_IDX(S.hot.path[d], c.top);
            const double sv = ((depth - 1 - d) & 1) ? -value : value;
            atomicAdd(&p->N, 1);
            atomicAdd(&p->W, sv);
        }
        if (lane == 0) S.hot.root_N += 1;  // every path starts at the root
        gsync();
    }

    // MCTS._simulate's descent (MCTS_model.py:372-392) with _select_child /
    // _get_ucb_score (:362-370, :129-139).  Returns 0 = reached a leaf,
    // 1 = reached a terminal node, -1 = error.  The parent's own virtual
    // visit is the "+1" under the square root; children carry none.
    __device__ int descend(int& leaf, int& depth, Node& nd)
    {
        int cur = c.root;
        nd.N = S.hot.root_N;  // root header from the mirror: no record load on the critical path
        nd.first_child = S.hot.root_fc;
        nd.meta = S.hot.root_meta;
        nd.moves = 0;
        depth = 0;
        const double cp64 = P.cfg.c_puct;
        const float cp32 = P.c_puct_f32;
        for (;;) {
            if (depth >= P.cfg.path_cap) {
                fail(OTH_ERR_PATH_OVERFLOW);
                return -1;
            }
            if (lane == 0) S.hot.path[depth] = cur;
            depth++;
            leaf = cur;
            if (meta_flags(nd.meta) & kFlagTerminal) return 1;
            if (nd.first_child < 0) return 0;
            const int fc = nd.first_child, nchild = meta_nchild(nd.meta);
            const double sq = sqrt((double)(nd.N + 1) + 1e-8);
            const float sq32 = (float)sq;
            const bool f64 = (depth == 1) && (c.flags & 2);
            int bi = 0x7fffffff;
            int k_N = 0, k_fc = -1;
            uint32_t k_meta = 0;
            u64 k_moves = 0;
            if (!f64) {
                float bs = -INFINITY;
                for (int i = lane; i < nchild; i += LANES) {
                    const Node ch = load_node(N + OTH_IDX(fc + i, c.top));
                    const double q = ch.N ? -__ddiv_rn(ch.W, (double)ch.N) : -0.0;
                    float u = __fmul_rn(cp32, ch.prior);
                    u = __fmul_rn(u, sq32);
                    u = __fdiv_rn(u, (float)(1 + ch.N));
                    const float sc = __fadd_rn((float)q, u);
                    if (sc > bs) {  // ascending i: the first maximum wins (:366-369)
                        bs = sc;
                        bi = i;
                        k_N = ch.N;
                        k_fc = ch.first_child;
                        k_meta = ch.meta;
                        k_moves = ch.moves;
                    }
                }
                const float m = redux_max_f32(bs, gmask);
                bi = (int)__reduce_min_sync(gmask, (unsigned)((bs == m) ? bi : 0x7fffffff));
            } else {
                double bs = -INFINITY;
                for (int i = lane; i < nchild; i += LANES) {
                    const Node ch = load_node(N + OTH_IDX(fc + i, c.top));
                    const double q = ch.N ? -__ddiv_rn(ch.W, (double)ch.N) : -0.0;
                    double u = __dmul_rn(cp64, P.root_prior64[(size_t)slot * OTH_MAX_CHILDREN + i]);
                    u = __dmul_rn(u, sq);
                    u = __ddiv_rn(u, (double)(1 + ch.N));
                    const double sc = __dadd_rn(q, u);
                    if (sc > bs) {
                        bs = sc;
                        bi = i;
                        k_N = ch.N;
                        k_fc = ch.first_child;
                        k_meta = ch.meta;
                        k_moves = ch.moves;
                    }
                }
#pragma unroll
                for (int o = LANES / 2; o; o >>= 1) {
                    const double os = gshfl_xor(bs, o);
                    const int oi = gshfl_xor(bi, o);
                    if (os > bs || (os == bs && oi < bi)) {
                        bs = os;
                        bi = oi;
                    }
                }
            }
            if ((unsigned)bi >= (unsigned)nchild) {  // no child compared greater than -inf: every score was NaN.  Cannot
                fail(OTH_ERR_NONFINITE);             // happen while expansion rejects non-finite priors / values; guards
                return -1;                           // the arena against an out-of-range child index all the same
            }
            const int src = bi & (LANES - 1);
            nd.N = gshfl(k_N, src);
            nd.first_child = gshfl(k_fc, src);
            nd.meta = gshfl(k_meta, src);
            if (nd.first_child < 0) nd.moves = gshfl(k_moves, src);  // a leaf: its legal set is what expansion needs
            cur = fc + bi;
            // the chosen child's board is needed only if it turns out to be the leaf: start fetching it now
            prefetch_l2(B + cur);
            count(OTH_CNT_LEVELS);
            count(OTH_CNT_CHILDREN, nchild);
        }
    }

    // --------------------------------------------------------- new game --
    __device__ void init_tree(u64 own, u64 opp, int player)
    {
        c.arena = 0;
        bind_arena();
        if (lane == 0) {
            Node r;
            r.W = 0.0;
            r.prior = 0.0f;
            r.N = 0;
            r.first_child = -1;
            r.moves = legal_moves(own, opp);
            r.meta = make_meta(0, 0xff, 0u, 0);  // a root is never terminal (MCTS_model.py:103-106)
            store_node(N, r);
            B[0] = make_ulonglong2(own, opp);
        }
        c.root = 0;
        c.top = 1;
        c.ply = 0;
        c.player = player;
        c.sims_done = 0;
        c.pending = -1;
        c.path_len = 0;
        c.flags = 0;
        c.reserved = -1;
        c.phase = OTH_PH_RUN;
        if (lane == 0) {
            S.hot.root_N = 0;
            S.hot.root_fc = -1;
            S.hot.root_meta = make_meta(0, 0xff, 0u, 0);
        }
        gsync();
        if (P.cfg.dirichlet_epsilon > 0.0) make_noise();  // this tree's root noise (kept out of the hot kernel)
    }

    // ---------------------------------------------------------- re-root --
    // MCTS.make_move (MCTS_model.py:200-215): keep the chosen subtree with its
    // statistics.  Breadth-first copy into the other arena; the destination
    // doubles as the BFS queue.
    __device__ void reroot(int new_root)
    {
        Node* Ns = N;
        ulonglong2* Bs = B;
        const int old_top = c.top;
        (void)old_top;
        c.arena ^= 1;
        bind_arena();
        Node* Nd = N;
        ulonglong2* Bd = B;
        if (lane == 0) {
            const Node r = load_node(Ns + new_root);
            store_node(Nd, r);
            Bd[0] = Bs[new_root];
        }
        gsync();
        // Breadth-first: the destination arena is the queue.  Each pass takes up to PER*LANES queued
        // nodes (all their headers are requested together), lays their child blocks out back to
        // back, then copies all those children as one flat list with PER copies in flight per lane --
        // two dependent round trips per pass instead of one per expanded node.
        constexpr int PER = LANES == 32 ? 2 : 4, CH = LANES * PER;
        static_assert((2 * CH + 1) * sizeof(int) <= sizeof(S.pri64), "re-root scratch must fit in pri64");
        int* s_ofc = reinterpret_cast<int*>(S.pri64);  // [CH] old first_child per queued node (scratch is free here)
        int* s_exc = s_ofc + CH;                        // [CH+1] exclusive prefix of child counts
        int top = 1, i = 0;
        while (i < top) {
            const int chunk = min(CH, top - i);
            int ofc[PER], nch[PER];
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const int k = j * LANES + lane;
                ofc[j] = -1;
                nch[j] = 0;
                if (k < chunk) {
                    const uint4 b = ldcg4(reinterpret_cast<const uint4*>(Nd + i + k) + 1);
                    ofc[j] = (int)b.x;
                    nch[j] = ofc[j] >= 0 ? (int)(b.y & 0xffu) : 0;
                }
            }
            int base = 0;
#pragma unroll
            for (int j = 0; j < PER; j++) {
                int incl = nch[j];
#pragma unroll
                for (int o = 1; o < LANES; o <<= 1) {
                    const int t = gshfl_up(incl, o);
                    if (lane >= o) incl += t;
                }
                const int excl = base + incl - nch[j];
                const int k = j * LANES + lane;
                s_ofc[k] = ofc[j];
                s_exc[k] = excl;
                if (k < chunk && nch[j] > 0) Nd[i + k].first_child = top + excl;
                base += gshfl(incl, LANES - 1);
            }
            const int total = base;
            gsync();
            for (int t0 = 0; t0 < total; t0 += CH) {
                uint4 x0[PER], x1[PER];
                ulonglong2 bb[PER];
                int dst[PER];
#pragma unroll
                for (int j = 0; j < PER; j++) {
                    const int t = t0 + j * LANES + lane;
                    dst[j] = -1;
                    if (t < total) {
                        int lo = 0, hi = chunk;  // last queued node whose child range starts at or before t
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (s_exc[mid] <= t) lo = mid;
                            else hi = mid;
                        }
                        const int src = s_ofc[lo] + (t - s_exc[lo]);
                        const uint4* sp = reinterpret_cast<const uint4*>(Ns + OTH_IDX(src, old_top));
                        x0[j] = ldcg4(sp);
                        x1[j] = ldcg4(sp + 1);
                        bb[j] = Bs[src];
                        dst[j] = top + t;
                    }
                }
#pragma unroll
                for (int j = 0; j < PER; j++) {
                    if (dst[j] >= 0) {
                        uint4* d = reinterpret_cast<uint4*>(Nd + OTH_IDX(dst[j], P.cfg.node_cap));
                        d[0] = x0[j];
                        d[1] = x1[j];
                        Bd[dst[j]] = bb[j];
                    }
                }
            }
            top += total;
            i += chunk;
            gsync();
        }
        count(OTH_CNT_COPIED, top);
        c.root = 0;
        c.top = top;
        c.flags &= ~2;
        if (lane == 0) {  // root mirror of the new root
            const Node r = load_node(Nd);
            S.hot.root_N = r.N;
            S.hot.root_fc = r.first_child;
            S.hot.root_meta = r.meta;
        }
        gsync();
    }

    // ------------------------------------------------ policy target etc --
    // Visit counts -> policy target in S.pri (MCTS_model.py:244-274).
    __device__ void policy_target(const Node& root, double temp, double u_tie)
    {
        const int fc = root.first_child, nchild = meta_nchild(root.meta);
        for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) S.pri[a] = 0.0f;
        gsync();
        for (int i = lane; i < nchild; i += LANES) {
            const Node ch = load_node(N + fc + i);
            S.pri[meta_action(ch.meta)] = (float)ch.N;
        }
        gsync();
        if (fabs(temp) < 1e-1) {
            if (lane == 0) {
                float mx = S.pri[0];
                for (int a = 1; a < OTH_NUM_ACTIONS; a++) mx = fmaxf(mx, S.pri[a]);
                int nt = 0;
                for (int a = 0; a < OTH_NUM_ACTIONS; a++) nt += (S.pri[a] == mx);
                int k = (int)floor(u_tie * (double)nt);
                if (k >= nt) k = nt - 1;
                int pick = 0;
                for (int a = 0, seen = 0; a < OTH_NUM_ACTIONS; a++)
                    if (S.pri[a] == mx) {
                        if (seen == k) pick = a;
                        seen++;
                    }
                for (int a = 0; a < OTH_NUM_ACTIONS; a++) S.pri[a] = (a == pick) ? 1.0f : 0.0f;
            }
            gsync();
            return;
        }
        if (temp != 1.0) {
            const float ex = (float)(1.0 / temp);
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) S.pri[a] = powf(S.pri[a], ex);
            gsync();
        }
        const float norm = np_sum65_f32();
        gsync();
        if (norm < (float)1e-12) {  // all-zero counts: uniform over valid actions (:260-269)
            const int nv = max(1, nchild);
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) S.pri[a] = 0.0f;
            gsync();
            for (int i = lane; i < nchild; i += LANES) {
                const Node ch = load_node(N + fc + i);
                S.pri[meta_action(ch.meta)] = (float)(1.0 / (double)nv);
            }
        } else {
            for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) S.pri[a] = __fdiv_rn(S.pri[a], norm);
        }
        gsync();
    }

    // np.random.choice(65, p) given its uniform (self_play_worker.py:75)
    __device__ int sample_action(double u)
    {
        int act = 0;
        if (lane == 0) {
            double last = 0.0;
            for (int a = 0; a < OTH_NUM_ACTIONS; a++) last = __dadd_rn(last, (double)S.pri[a]);
            double acc = 0.0;
            for (int a = 0; a < OTH_NUM_ACTIONS; a++) {
                acc = __dadd_rn(acc, (double)S.pri[a]);
                if (__ddiv_rn(acc, last) <= u) act = a + 1;
            }
        }
        return gshfl(act, 0);
    }

    __device__ double uniform_for(uint32_t purpose, int ply)
    {
        const Philox4 r = philox4x32_10(P.cfg.seed, (uint64_t)c.game_id, (uint32_t)ply, purpose);
        return u01_53(r.x, r.y);
    }

    // get_training_data (self_play_worker.py:8-35) + hand-off of the game's
    // replay tuples to the output ring.
    __device__ void emit_game(int nply, int winner)
    {
        long long base = 0, gi = 0;
        if (lane == 0) {
            base = (long long)atomicAdd(P.counters + OTH_CNT_POSITIONS, (unsigned long long)nply);
            gi = (long long)atomicAdd(P.counters + OTH_CNT_OUT_GAMES, 1ULL);
        }
        base = gshfl(base, 0);
        gi = gshfl(gi, 0);
        if (base + nply > P.cfg.out_pos_cap || gi >= P.cfg.out_game_cap) {
            fail(OTH_ERR_OUT_OVERFLOW);
            return;
        }
        const size_t tb = (size_t)slot * OTH_MAX_PLIES;
        if (lane == 0) {
            const double lam = P.cfg.lambda;
            double g_next = 0.0;
            int next_player = 0;
            for (int t = nply - 1; t >= 0; t--) {
                const int pl = (int)(int8_t)(P.traj_meta[tb + t] & 0xff);
                const double mc = winner == 0 ? 0.0 : (pl == winner ? 1.0 : -1.0);
                double g;
                if (t == nply - 1) g = mc;
                else {
                    const double sign = pl == next_player ? 1.0 : -1.0;
                    const double a = __dmul_rn(1.0 - lam, P.traj_rootv[tb + t]);
                    const double b = __dmul_rn(__dmul_rn(lam, sign), g_next);
                    g = __dadd_rn(a, b);
                }
                P.out_value[base + t] = g;
                g_next = g;
                next_player = pl;
            }
            long long* gd = P.out_games + gi * 4;
            gd[0] = c.game_id;
            gd[1] = base;
            gd[2] = nply;
            gd[3] = winner;
        }
        for (int t = lane; t < nply; t += LANES) {
            P.out_board[base + t] = P.traj_board[tb + t];
            const int m = P.traj_meta[tb + t];
            P.out_meta[base + t] = (c.game_id << 16) | ((long long)t << 8) | (long long)(m & 0xff);
        }
        const float* sp = P.traj_pi + tb * OTH_NUM_ACTIONS;
        float* dp = P.out_pi + (size_t)base * OTH_NUM_ACTIONS;
        for (int e = lane; e < nply * OTH_NUM_ACTIONS; e += LANES) dp[e] = sp[e];
        count(OTH_CNT_GAMES);
    }

    // One self-play ply after its search (self_play_worker.py:64-88).
    __device__ void finish_move()
    {
        const Node root = load_node(N + c.root);
        const int T = c.ply;
        if (T >= OTH_MAX_PLIES) {
            fail(OTH_ERR_PLY_OVERFLOW);
            return;
        }
        const size_t tb = (size_t)slot * OTH_MAX_PLIES;
        const double temp = T < P.cfg.num_exploratory_moves ? P.cfg.temperature : 0.0;
        double ut = 0.0, um = 0.0;
        if (P.cfg.inject_random) {
            ut = P.u_tie[tb + T];
            um = P.u_move[tb + T];
        } else {
            ut = uniform_for(kPurposeTie, T);
            um = uniform_for(kPurposeMove, T);
            if (lane == 0) {
                P.u_tie[tb + T] = ut;
                P.u_move[tb + T] = um;
            }
        }
        policy_target(root, temp, ut);
        for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) P.traj_pi[(tb + T) * OTH_NUM_ACTIONS + a] = S.pri[a];
        const int action = sample_action(um);
        if (lane == 0) {
            P.traj_board[tb + T] = B[c.root];
            P.traj_rootv[tb + T] = root.N ? __ddiv_rn(root.W, (double)root.N) : 0.0;  // mcts.root.value, :72-73
            P.traj_meta[tb + T] = (c.player & 0xff) | (action << 8);
        }
        count(OTH_CNT_MOVES);
        // mcts.make_move(action): locate the child (KeyError if absent)
        const int fc = root.first_child, nchild = meta_nchild(root.meta);
        int ci = -1;
        uint32_t cmeta = 0;
        for (int i = lane; i < nchild; i += LANES) {
            const uint4 b = ldcg4(reinterpret_cast<const uint4*>(N + fc + i) + 1);
            if (meta_action(b.y) == action) {
                ci = i;
                cmeta = b.y;
            }
        }
        const unsigned hit = gballot(ci >= 0);
        if (!hit) {
            fail(OTH_ERR_BAD_ACTION);
            return;
        }
        const int src = __ffs(hit) - 1;
        ci = gshfl(ci, src);
        cmeta = gshfl(cmeta, src);
        gsync();
        if (meta_flags(cmeta) & kFlagTerminal) {
            // reward is read from the mover's side (:78-82); the child's terminal value is the opponent's
            const int reward = -meta_tvalue(cmeta);
            const int winner = reward > 0 ? c.player : (reward < 0 ? -c.player : 0);
            emit_game(T + 1, winner);
            if (c.phase == OTH_PH_ERROR) return;
            if (c.games_left > 0) c.games_left--;
            if (c.games_left == 0) {
                c.phase = OTH_PH_DONE;
                return;
            }
            c.game_id += (long long)P.cfg.game_id_stride;
            init_tree(INIT_BLACK, INIT_WHITE, 1);
            return;
        }
        reroot(fc + ci);
        c.ply = T + 1;
        c.player = -c.player;
        c.sims_done = 0;
    }

    // ------------------------------------------------------ slot driver --
    // MOVE: this instantiation may finish moves (policy target, sampling, re-rooting, game
    //       hand-off).  The hot kernel (MOVE = false) only flags such slots; k_mcts_move picks
    //       them up in the same oth_mcts_step call with a full warp per slot.
    // STUB: device evaluators compiled in (search-only / test builds of the kernel).
    // FUSED: the network's softmax (Models.py:24-25) and tanh are applied here, on raw logits.
    // DEVSTUB: the production split (hot kernel + move kernel) with a device stub standing in for the
    //       network: the pending leaf is evaluated by the stub when the hot kernel consumes it, so a
    //       launch does exactly what it does behind the network -- one evaluation per slot.
    template <bool MOVE, bool STUB, bool FUSED = false, bool DEVSTUB = false>
    __device__ void run_slot()
    {
        // (1) everything that depends only on the slot index is requested at once:
        //     control block, network outputs for the pending leaf
        constexpr int NPL = (OTH_NUM_ACTIONS + LANES - 1) / LANES;
        constexpr bool stub = STUB;
        float pv[NPL];
        float nn_value = 0.0f;
        int eval_row = slot;
        if (!stub && !DEVSTUB && !(MOVE && !STUB)) {
            if constexpr (FUSED) {
                if (P.eval_map) eval_row = P.eval_map[slot];
                const size_t row = (size_t)(eval_row < 0 ? 0 : eval_row);  // a slot without a row reads row 0 and discards it
                if (P.raw_bf16) {
                    const __nv_bfloat16* lg = (const __nv_bfloat16*)P.logits + row * P.logits_stride;
#pragma unroll
                    for (int k = 0; k < NPL; k++) {
                        const int a = lane + k * LANES;
                        pv[k] = a < OTH_NUM_ACTIONS ? __bfloat162float(lg[a]) : -INFINITY;
                    }
                    nn_value = __bfloat162float(((const __nv_bfloat16*)P.vpre)[row * P.vpre_stride]);
                } else {
                    const float* lg = (const float*)P.logits + row * P.logits_stride;
#pragma unroll
                    for (int k = 0; k < NPL; k++) {
                        const int a = lane + k * LANES;
                        pv[k] = a < OTH_NUM_ACTIONS ? lg[a] : -INFINITY;
                    }
                    nn_value = ((const float*)P.vpre)[row * P.vpre_stride];
                }
            } else {
                const float* pr = P.priors + (size_t)slot * OTH_NUM_ACTIONS;
#pragma unroll
                for (int k = 0; k < NPL; k++) {
                    const int a = lane + k * LANES;
                    pv[k] = a < OTH_NUM_ACTIONS ? pr[a] : 0.0f;
                }
                nn_value = P.values[slot];
            }
        }
        load_counters();
        load_hot_head();
        c = P.ctl[slot];
        if (c.top < 1) return;  // slot never given a tree (zero-filled control block): nothing to do
        if (FUSED && eval_row < 0 && c.phase == OTH_PH_WAIT_EVAL) return;  // its position did not fit this launch's bucket: it waits
        bind_arena();
        gsync();
        if (c.phase == OTH_PH_WAIT_EVAL || c.phase == OTH_PH_RUN) {
            // the next descent starts at the root's children: get them into L2 now
            const int rfc = S.hot.root_fc, rn = meta_nchild(S.hot.root_meta);
            if (rfc >= 0)
                for (int i = lane; i < rn; i += LANES) prefetch_l2(N + rfc + i);
        }
        if constexpr (MOVE) {
            if (c.phase == OTH_PH_MOVE) {
                c.phase = OTH_PH_RUN;
                finish_move();
            }
        }
        if (!(MOVE && !STUB) && c.phase == OTH_PH_WAIT_EVAL) {
            // (2) no second round trip: the pending leaf's board / legal set / meta word and the
            //     path came with the hot record (only paths deeper than 8 need more of it)
            load_path_rest(c.path_len);
            if constexpr (FUSED) {
                // softmax over the 65 logits: max, exp, sum (lanes combined in a fixed order), divide
                float mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < NPL; k++) mx = fmaxf(mx, pv[k]);
                mx = redux_max_f32(mx, gmask);
                float sum = 0.0f;
#pragma unroll
                for (int k = 0; k < NPL; k++) {
                    const int a = lane + k * LANES;
                    pv[k] = a < OTH_NUM_ACTIONS ? expf(pv[k] - mx) : 0.0f;
                    sum = __fadd_rn(sum, pv[k]);
                }
#pragma unroll
                for (int o = LANES / 2; o; o >>= 1) sum = __fadd_rn(sum, __shfl_down_sync(gmask, sum, o, LANES));
                sum = gshfl(sum, 0);
#pragma unroll
                for (int k = 0; k < NPL; k++) pv[k] = __fdiv_rn(pv[k], sum);
                nn_value = tanhf(nn_value);
                if (P.priors_out) {
#pragma unroll
                    for (int k = 0; k < NPL; k++) {
                        const int a = lane + k * LANES;
                        if (a < OTH_NUM_ACTIONS) P.priors_out[(size_t)slot * OTH_NUM_ACTIONS + a] = pv[k];
                    }
                    if (lane == 0) P.values_out[slot] = nn_value;
                }
            }
            Node lf;
            lf.moves = S.hot.leaf_moves;
            lf.meta = S.hot.leaf_meta;
            lf.first_child = -1;
            const ulonglong2 lb = make_ulonglong2(S.hot.leaf_own, S.hot.leaf_opp);
            double leaf_value;
            if constexpr (DEVSTUB) {
                leaf_value = eval_stub(lb.x, lb.y);  // raw priors to S.pri; masked inside expand
            } else {
#pragma unroll
                for (int k = 0; k < NPL; k++) {  // priors *= valid_mask (:346) fused into the staging store
                    const int a = lane + k * LANES;
                    const bool valid = a < 64 ? ((lf.moves >> a) & 1) : (lf.moves == 0);
                    if (a < OTH_NUM_ACTIONS) S.pri[a] = valid ? pv[k] : __fmul_rn(pv[k], 0.0f);
                }
                gsync();
                leaf_value = (double)nn_value;
            }
            const bool root_init = c.flags & 1;
            if (!isfinite(leaf_value)) fail(OTH_ERR_NONFINITE);
            else if (expand(c.pending, root_init, lf, lb, !DEVSTUB)) {
                backup(c.path_len, leaf_value);
                if (!root_init) {
                    c.sims_done++;
                    count(OTH_CNT_SIMS);
                }
                count(OTH_CNT_EVALS);
                count_max(OTH_CNT_MAX_DEPTH, (unsigned)c.path_len);
                c.flags &= ~1;
                c.pending = -1;
                c.phase = OTH_PH_RUN;
            }
        }
        int budget = P.cfg.max_inline_sims;
        while (c.phase == OTH_PH_RUN && budget > 0) {
            if (c.sims_done >= P.cfg.num_simulations) {
                if (!P.cfg.self_play) {
                    c.phase = OTH_PH_IDLE;
                    break;
                }
                if constexpr (MOVE) {
                    finish_move();
                    budget--;
                    continue;
                } else {
                    c.phase = OTH_PH_MOVE;  // the move kernel of this same step takes over
                    if (lane == 0) {
                        P.move_flags[slot] = 1;
                        if (P.cfg.move_launch) {
                            P.move_list[4 + atomicAdd(P.move_list, 1)] = slot;
                        }
                    }
                    break;
                }
            }
            int leaf, depth;
            Node nd;
            const int r = descend(leaf, depth, nd);
            if (r < 0) break;
            if (r == 1) {  // terminal node: back up its value again, no evaluation (:381-384)
                backup(depth, (double)meta_tvalue(nd.meta));
                c.sims_done++;
                budget--;
                count(OTH_CNT_SIMS);
                count(OTH_CNT_TERMINAL);
                continue;
            }
            const bool root_init = (depth == 1);  // the leaf is the root: policy_improve_step :234-235
            const ulonglong2 lb = B[OTH_IDX(leaf, c.top)];
            if (root_init) {  // the mirror has no legal set: read the root record (once per game at most)
                const Node rr = load_node(N + leaf);
                nd.moves = rr.moves;
                nd.meta = rr.meta;
            }
            if constexpr (STUB) {
                const double value = P.cfg.eval_kind == OTH_EVAL_ROLLOUT ? eval_rollout(lb.x, lb.y, root_init) : eval_stub(lb.x, lb.y);
                if (!expand(leaf, root_init, nd, lb, false)) break;
                backup(depth, value);
                if (!root_init) {
                    c.sims_done++;
                    count(OTH_CNT_SIMS);
                }
                count(OTH_CNT_EVALS);
                count_max(OTH_CNT_MAX_DEPTH, (unsigned)depth);
                budget--;
                continue;
            }
            c.pending = leaf;
            c.path_len = depth;
            c.flags = (c.flags & ~1) | (root_init ? 1 : 0);
            c.phase = OTH_PH_WAIT_EVAL;
            if (!DEVSTUB || P.nn_input) write_nn_input(lb.x, lb.y);
            if (lane == 0) {  // what the expansion in the next launch needs, carried in the hot record
                S.hot.leaf_own = lb.x;
                S.hot.leaf_opp = lb.y;
                S.hot.leaf_moves = nd.moves;
                S.hot.leaf_meta = nd.meta;
            }
        }
        count_max(OTH_CNT_MAX_TOP, (unsigned)c.top);  // arena high-water mark of this slot
#ifdef OTH_DEBUG
        {  // an assertion may have fired on any lane: lane 0 writes the control block
            int e = c.error;
            long long r = (c.error & OTH_ERR_DEBUG) ? c.reserved : -1;
#pragma unroll
            for (int o = LANES / 2; o; o >>= 1) {
                e |= gshfl_xor(e, o);
                r = max(r, gshfl_xor(r, o));
            }
            if (e & OTH_ERR_DEBUG) {
                c.error = e;
                c.reserved = r;
                c.phase = OTH_PH_ERROR;
            }
        }
#endif
        if (lane == 0) P.ctl[slot] = c;
        store_hot(c.phase == OTH_PH_WAIT_EVAL ? c.path_len : 0);
        store_counters();
        gsync();
    }
};

// The hot kernel (external network): expansion, backup, descent.  Moves are only flagged.
template <int LANES>
__global__ void __launch_bounds__(kBlock, kStepBlocksPerSM) k_mcts_step(const Params P)
{
    __shared__ Scratch scratch[kBlock / LANES];
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    Ctx<LANES> ctx(tile, P, scratch[threadIdx.x / LANES]);
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        ctx.slot = s;
        ctx.template run_slot<false, false>();
    }
}

// The hot kernel with a device stub in the network's place (cfg.split_stub): search-only runs and parity tests of
// the production kernel pair at sizes where recording a real network's outputs is impractical.
template <int LANES>
__global__ void __launch_bounds__(kBlock, kStepBlocksPerSM) k_mcts_step_devstub(const Params P)
{
    __shared__ Scratch scratch[kBlock / LANES];
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    Ctx<LANES> ctx(tile, P, scratch[threadIdx.x / LANES]);
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        ctx.slot = s;
        ctx.template run_slot<false, false, false, true>();
    }
}

// The hot kernel with the network's softmax / tanh fused in (oth_mcts_step_fused).
template <int LANES>
__global__ void __launch_bounds__(kBlock, kStepBlocksPerSM) k_mcts_step_fused(const Params P)
{
    __shared__ Scratch scratch[kBlock / LANES];
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    Ctx<LANES> ctx(tile, P, scratch[threadIdx.x / LANES]);
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        ctx.slot = s;
        ctx.template run_slot<false, false, true>();
    }
}

// Device-evaluator build (stubs / rollouts): everything in one kernel, whole simulations per launch.
template <int LANES>
__global__ void __launch_bounds__(kBlock, 4) k_mcts_step_stub(const Params P)
{
    __shared__ Scratch scratch[kBlock / LANES];
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    Ctx<LANES> ctx(tile, P, scratch[threadIdx.x / LANES]);
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        ctx.slot = s;
        ctx.template run_slot<true, true>();
    }
}

// Move kernel: one full warp per flagged slot -- policy target, move sampling, trajectory row,
// breadth-first re-rooting or game hand-off and restart, then the descent that puts the slot's
// next leaf into the network batch, so a slot never misses an iteration.
__global__ void __launch_bounds__(kBlock, 4) k_mcts_move(const Params P)
{
    __shared__ Scratch scratch[kBlock / 32];
    cg::thread_block_tile<32> tile = cg::tiled_partition<32>(cg::this_thread_block());
    Ctx<32> ctx(tile, P, scratch[threadIdx.x / 32]);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * kBlock) / 32;
    // a warp owns kMoveSet consecutive slots: small sets so that the launch in which every slot re-roots at once
    // (one per move while the games are in step) spreads over ~2 000 warps instead of 512
    const int n_sets = (P.cfg.n_slots + kMoveSet - 1) / kMoveSet;
    for (int w = (blockIdx.x * kBlock + threadIdx.x) / 32; w < n_sets; w += warps) {
        const int s0 = w * kMoveSet + lane;
        unsigned m = __ballot_sync(0xffffffffu, lane < kMoveSet && s0 < P.cfg.n_slots && P.move_flags[s0] != 0);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            ctx.slot = w * kMoveSet + b;
            ctx.template run_slot<true, false>();
            if (lane == 0) P.move_flags[ctx.slot] = 0;
            __syncwarp();
        }
    }
}

// The move kernel, due-list form (cfg.move_launch = 1): the step kernel appended every slot whose move is due to
// OTH_BUF_MOVE_LIST.  A block reads the count first and exits at once when it is zero -- the common launch, in which no
// slot moves, costs one 4-byte load per block instead of a scan of the move flags -- otherwise one warp takes one due
// slot (no scan, and the launch in which every slot moves is spread evenly).  The last block to finish resets the list.
// (A device-side tail launch of this kernel from the step kernel was tried and removed: inside a captured CUDA graph the
// child grid is not ordered before the next graph launch, so its control-block updates raced with the next step kernel.)
// (4 resident blocks per SM = 128 registers: the launch in which all 16 384 slots re-root is bound by how many BFS copies are
// in flight -- 394 us at 128 registers, 615 us unconstrained at 162, 471 us at 96, 532 us at 80 with their spills.)
__global__ void __launch_bounds__(kBlock, 4) k_mcts_move_list(const Params P)
{
    __shared__ Scratch scratch[kBlock / 32];
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = *reinterpret_cast<volatile int*>(P.move_list);
    __syncthreads();
    const int n_due = s_n;
    if (n_due == 0) return;
    cg::thread_block_tile<32> tile = cg::tiled_partition<32>(cg::this_thread_block());
    Ctx<32> ctx(tile, P, scratch[threadIdx.x / 32]);
    const int warps = (gridDim.x * kBlock) / 32;
    for (int w = (blockIdx.x * kBlock + threadIdx.x) / 32; w < n_due; w += warps) {
        ctx.slot = P.move_list[4 + w];
        ctx.template run_slot<true, false>();
        if ((threadIdx.x & 31) == 0) P.move_flags[ctx.slot] = 0;
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(P.move_list + 1, 1) == (int)gridDim.x - 1) {  // every block has read the count: reset for the next step kernel
            P.move_list[0] = 0;
            P.move_list[1] = 0;
        }
    }
}

// Counters on demand (kept out of the hot kernel): sums the per-slot event counters and derives
// the gauges (slots waiting for the network, slots still playing, slots in error, fullest arena)
// from the control blocks.  POSITIONS / OUT_GAMES are live atomics of emit_game and left alone.
__global__ void __launch_bounds__(256) k_mcts_poll(const Params P)
{
    __shared__ unsigned long long acc[CNT_LOCAL];
    if (threadIdx.x < CNT_LOCAL) acc[threadIdx.x] = 0;
    __syncthreads();
    const int i = threadIdx.x & (CNT_LOCAL - 1);
    const bool is_max = (i == OTH_CNT_MAX_TOP || i == OTH_CNT_MAX_DEPTH);
    unsigned long long v = 0;
    // 16 consecutive threads read one slot's 16 counters (64 B)
    for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / CNT_LOCAL; s < P.cfg.n_slots;
         s += ((long long)gridDim.x * blockDim.x) / CNT_LOCAL) {
        unsigned x = P.slot_counters[s * CNT_LOCAL + i];
        const oth_mcts_ctl* c = P.ctl + s;
        if (i == OTH_CNT_WAITING) x = c->phase == OTH_PH_WAIT_EVAL;
        if (i == OTH_CNT_ACTIVE) x = c->top >= 1 && (c->phase == OTH_PH_WAIT_EVAL || c->phase == OTH_PH_RUN || c->phase == OTH_PH_MOVE);
        if (i == OTH_CNT_ERRORS) x = (c->phase == OTH_PH_ERROR || c->error != 0);
        v = is_max ? (x > v ? x : v) : v + x;
    }
    if (v) {
        if (is_max) atomicMax(acc + i, v);
        else atomicAdd(acc + i, v);
    }
    __syncthreads();
    if (threadIdx.x < CNT_LOCAL && acc[threadIdx.x] && i != OTH_CNT_POSITIONS && i != OTH_CNT_OUT_GAMES) {
        if (is_max) atomicMax(P.counters + i, acc[i]);
        else atomicAdd(P.counters + i, acc[i]);
    }
}

// Start fresh games (self-play) on every slot.
template <int LANES>
__global__ void __launch_bounds__(kBlock) k_mcts_reset(const Params P)
{
    __shared__ Scratch scratch[kBlock / LANES];
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    Ctx<LANES> ctx(tile, P, scratch[threadIdx.x / LANES]);
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        ctx.slot = s;
        oth_mcts_ctl z = {};
        ctx.c = z;
        ctx.c.game_id = (long long)(P.cfg.game_id_base + (uint64_t)s);
        ctx.c.games_left = P.cfg.games_per_slot < 0 ? -1 : P.cfg.games_per_slot;
        ctx.init_tree(INIT_BLACK, INIT_WHITE, 1);
        if (P.cfg.games_per_slot == 0) ctx.c.phase = OTH_PH_DONE;
        for (int i = ctx.lane; i < CNT_LOCAL; i += LANES) P.slot_counters[(size_t)s * CNT_LOCAL + i] = 0;
        if (ctx.lane == 0) P.ctl[s] = ctx.c;
        ctx.store_hot(0);
        tile.sync();
    }
}

template <int LANES>
__global__ void __launch_bounds__(kBlock) k_mcts_set_roots(const Params P, const u64* own, const u64* opp, const int8_t* players,
                                                           const uint8_t* mask)
{
    __shared__ Scratch scratch[kBlock / LANES];
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    Ctx<LANES> ctx(tile, P, scratch[threadIdx.x / LANES]);
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        if (mask && !mask[s]) continue;
        ctx.slot = s;
        oth_mcts_ctl z = {};
        ctx.c = z;
        ctx.c.game_id = (long long)(P.cfg.game_id_base + (uint64_t)s);
        ctx.c.games_left = -1;
        ctx.init_tree(own[s], opp[s], players[s]);
        ctx.c.phase = OTH_PH_IDLE;
        for (int i = ctx.lane; i < CNT_LOCAL; i += LANES) P.slot_counters[(size_t)s * CNT_LOCAL + i] = 0;
        if (ctx.lane == 0) P.ctl[s] = ctx.c;
        ctx.store_hot(0);
        tile.sync();
    }
}

__global__ void k_mcts_begin_search(const Params P, const uint8_t* mask)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.cfg.n_slots || (mask && !mask[s])) return;
    oth_mcts_ctl* c = P.ctl + s;
    if (c->phase == OTH_PH_IDLE && c->top >= 1) {
        c->sims_done = 0;
        c->phase = OTH_PH_RUN;
    }
}

template <int LANES>
__global__ void __launch_bounds__(kBlock) k_mcts_advance(const Params P, const int32_t* actions)
{
    __shared__ Scratch scratch[kBlock / LANES];
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    Ctx<LANES> ctx(tile, P, scratch[threadIdx.x / LANES]);
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        const int action = actions[s];
        if (action < 0) continue;
        ctx.slot = s;
        ctx.c = P.ctl[s];
        if (ctx.c.phase == OTH_PH_ERROR || ctx.c.top < 1) continue;
        ctx.load_counters();
        ctx.load_hot_head();
        ctx.bind_arena();
        tile.sync();
        const Node root = load_node(ctx.N + ctx.c.root);
        const int fc = root.first_child, nchild = root.first_child < 0 ? 0 : meta_nchild(root.meta);
        int ci = -1;
        for (int i = ctx.lane; i < nchild; i += LANES) {
            const uint4 b = ldcg4(reinterpret_cast<const uint4*>(ctx.N + fc + i) + 1);
            if (meta_action(b.y) == action) ci = i;
        }
        const unsigned hit = tile.ballot(ci >= 0);
        if (!hit) {
            ctx.c.error |= OTH_ERR_BAD_ACTION;  // KeyError; the tree is left as it was
        } else {
            ci = tile.shfl(ci, __ffs(hit) - 1);
            ctx.reroot(fc + ci);
            ctx.c.ply += 1;
            ctx.c.player = -ctx.c.player;
            ctx.c.sims_done = 0;
            ctx.c.phase = OTH_PH_IDLE;
            // a search that starts on an unexpanded root draws fresh Dirichlet noise (MCTS_model.py:234-235, 339-343)
            if (ctx.S.hot.root_fc < 0 && P.cfg.dirichlet_epsilon > 0.0) ctx.make_noise();
        }
        if (ctx.lane == 0) P.ctl[s] = ctx.c;
        ctx.store_hot(0);
        ctx.store_counters();
        tile.sync();
    }
}

template <int LANES>
__global__ void __launch_bounds__(kBlock) k_mcts_root_stats(const Params P, int32_t* counts, double* child_value, double* child_prior,
                                                            double* root_value, int32_t* root_n, u64* root_board)
{
    cg::thread_block_tile<LANES> tile = cg::tiled_partition<LANES>(cg::this_thread_block());
    const int lane = tile.thread_rank();
    const int groups = (gridDim.x * kBlock) / LANES;
    for (int s = (blockIdx.x * kBlock + threadIdx.x) / LANES; s < P.cfg.n_slots; s += groups) {
        const oth_mcts_ctl c = P.ctl[s];
        const size_t base = ((size_t)s * 2 + (size_t)c.arena) * (size_t)P.cfg.node_cap;
        const Node* N = P.nodes + base;
        const Node root = load_node(N + c.root);
        for (int a = lane; a < OTH_NUM_ACTIONS; a += LANES) {
            if (counts) counts[(size_t)s * OTH_NUM_ACTIONS + a] = 0;
            if (child_value) child_value[(size_t)s * OTH_NUM_ACTIONS + a] = 0.0;
            if (child_prior) child_prior[(size_t)s * OTH_NUM_ACTIONS + a] = 0.0;
        }
        tile.sync();
        const int nchild = root.first_child < 0 ? 0 : meta_nchild(root.meta);
        for (int i = lane; i < nchild; i += LANES) {
            const Node ch = load_node(N + root.first_child + i);
            const int a = meta_action(ch.meta);
            if (counts) counts[(size_t)s * OTH_NUM_ACTIONS + a] = ch.N;
            if (child_value) child_value[(size_t)s * OTH_NUM_ACTIONS + a] = ch.N ? ch.W / (double)ch.N : 0.0;
            if (child_prior)
                child_prior[(size_t)s * OTH_NUM_ACTIONS + a] =
                    (c.flags & 2) ? P.root_prior64[(size_t)s * OTH_MAX_CHILDREN + i] : (double)ch.prior;
        }
        if (lane == 0) {
            if (root_value) root_value[s] = root.N ? root.W / (double)root.N : 0.0;
            if (root_n) root_n[s] = root.N;
            if (root_board) {
                const ulonglong2 b = P.boards[base + c.root];
                root_board[2 * (size_t)s] = b.x;
                root_board[2 * (size_t)s + 1] = b.y;
            }
        }
    }
}

int check_cfg(const oth_mcts_config* cfg)
{
    if (!cfg) return OTH_E_ARG;
    if (cfg->n_slots <= 0 || cfg->node_cap < 64 || cfg->path_cap < 2 || cfg->path_cap > 128) return OTH_E_ARG;
    if (cfg->num_simulations < 0 || cfg->max_inline_sims <= 0) return OTH_E_ARG;
    if (cfg->lanes != 8 && cfg->lanes != 16 && cfg->lanes != 32) return OTH_E_ARG;
    if (cfg->eval_kind < OTH_EVAL_EXTERNAL || cfg->eval_kind > OTH_EVAL_ROLLOUT) return OTH_E_ARG;
    if (cfg->hot_path < 0 || cfg->hot_path > kHotPath) return OTH_E_ARG;
    if (cfg->split_stub && (cfg->eval_kind < OTH_EVAL_STUB_A || cfg->eval_kind > OTH_EVAL_STUB_H)) return OTH_E_ARG;
    if (cfg->move_launch != 0 && cfg->move_launch != 1) return OTH_E_ARG;
    if (cfg->self_play && (cfg->out_pos_cap <= 0 || cfg->out_game_cap <= 0)) return OTH_E_ARG;
    return OTH_OK;
}

int make_params(const oth_mcts_config* cfg, const oth_mcts_buffers* b, Params* p)
{
    const int rc = check_cfg(cfg);
    if (rc != OTH_OK) return rc;
    if (!b) return OTH_E_ARG;
    for (int i = 0; i < OTH_BUF_COUNT; i++) {
        const bool out_buf = i >= OTH_BUF_OUT_BOARD && i <= OTH_BUF_OUT_GAMES;
        const bool traj_buf = i >= OTH_BUF_TRAJ_BOARD && i <= OTH_BUF_TRAJ_META;
        if (!b->buf[i] && !((out_buf || traj_buf) && !cfg->self_play)) return OTH_E_ARG;
    }
    p->cfg = *cfg;
    p->nodes = (Node*)b->buf[OTH_BUF_NODES];
    p->boards = (ulonglong2*)b->buf[OTH_BUF_BOARDS];
    p->ctl = (oth_mcts_ctl*)b->buf[OTH_BUF_CTL];
    p->path = (int*)b->buf[OTH_BUF_PATH];
    p->root_prior64 = (double*)b->buf[OTH_BUF_ROOT_PRIOR64];
    p->noise = (double*)b->buf[OTH_BUF_NOISE];
    p->u_move = (double*)b->buf[OTH_BUF_U_MOVE];
    p->u_tie = (double*)b->buf[OTH_BUF_U_TIE];
    p->traj_board = (ulonglong2*)b->buf[OTH_BUF_TRAJ_BOARD];
    p->traj_pi = (float*)b->buf[OTH_BUF_TRAJ_PI];
    p->traj_rootv = (double*)b->buf[OTH_BUF_TRAJ_ROOTV];
    p->traj_meta = (int*)b->buf[OTH_BUF_TRAJ_META];
    p->out_board = (ulonglong2*)b->buf[OTH_BUF_OUT_BOARD];
    p->out_pi = (float*)b->buf[OTH_BUF_OUT_PI];
    p->out_value = (double*)b->buf[OTH_BUF_OUT_VALUE];
    p->out_meta = (long long*)b->buf[OTH_BUF_OUT_META];
    p->out_games = (long long*)b->buf[OTH_BUF_OUT_GAMES];
    p->counters = (unsigned long long*)b->buf[OTH_BUF_COUNTERS];
    p->slot_counters = (unsigned*)b->buf[OTH_BUF_SLOT_COUNTERS];
    p->hot = (uint4*)b->buf[OTH_BUF_HOT];
    p->move_flags = (uint8_t*)b->buf[OTH_BUF_MOVE_FLAGS];
    p->move_list = (int*)b->buf[OTH_BUF_MOVE_LIST];
    p->hot_path = cfg->hot_path > 0 ? cfg->hot_path : kHotPath;
    p->priors = nullptr;
    p->values = nullptr;
    p->logits = nullptr;
    p->vpre = nullptr;
    p->logits_stride = p->vpre_stride = p->raw_bf16 = 0;
    p->eval_map = nullptr;
    p->priors_out = nullptr;
    p->values_out = nullptr;
    p->nn_input = nullptr;
    p->c_puct_f32 = (float)cfg->c_puct;
    return OTH_OK;
}

// Persistent-style grid: whole SMs' worth of blocks, groups stride over the slots.
int mcts_grid(const oth_mcts_config* cfg)
{
    const int64_t need = ((int64_t)cfg->n_slots * cfg->lanes + kBlock - 1) / kBlock;
    const int64_t full = (int64_t)sm_count() * 16;  // 16 blocks of 128 threads fill an SM
    return (int)(need < full ? need : full);
}

// k_mcts_move: one warp per set of kMoveSet slots, 4 resident blocks per SM (128 registers)
inline int move_grid(const oth_mcts_config* cfg)
{
    const int warps = (cfg->n_slots + kMoveSet - 1) / kMoveSet;
    const int blocks = (warps + kBlock / 32 - 1) / (kBlock / 32);
    return blocks < sm_count() * 4 ? blocks : sm_count() * 4;
}

#define LAUNCH_LANES(kernel, grid, stream, ...)                                              \
    do {                                                                                     \
        if (cfg->lanes == 32) kernel<32><<<grid, kBlock, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
        else if (cfg->lanes == 16) kernel<16><<<grid, kBlock, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
        else kernel<8><<<grid, kBlock, 0, (cudaStream_t)stream>>>(__VA_ARGS__);                \
    } while (0)

}  // namespace

extern "C" int oth_mcts_buffer_bytes(const oth_mcts_config* cfg, int64_t* out)
{
    const int rc = check_cfg(cfg);
    if (rc != OTH_OK) return rc;
    if (!out) return OTH_E_ARG;
    const int64_t G = cfg->n_slots, cap = cfg->node_cap, T = OTH_MAX_PLIES;
    const bool sp = cfg->self_play != 0;
    out[OTH_BUF_NODES] = G * 2 * cap * 32;
    out[OTH_BUF_BOARDS] = G * 2 * cap * 16;
    out[OTH_BUF_CTL] = G * (int64_t)sizeof(oth_mcts_ctl);
    out[OTH_BUF_PATH] = G * cfg->path_cap * 4;
    out[OTH_BUF_ROOT_PRIOR64] = G * OTH_MAX_CHILDREN * 8;
    out[OTH_BUF_NOISE] = G * OTH_NUM_ACTIONS * 8;
    out[OTH_BUF_U_MOVE] = G * T * 8;
    out[OTH_BUF_U_TIE] = G * T * 8;
    out[OTH_BUF_TRAJ_BOARD] = sp ? G * T * 16 : 0;
    out[OTH_BUF_TRAJ_PI] = sp ? G * T * OTH_NUM_ACTIONS * 4 : 0;
    out[OTH_BUF_TRAJ_ROOTV] = sp ? G * T * 8 : 0;
    out[OTH_BUF_TRAJ_META] = sp ? G * T * 4 : 0;
    out[OTH_BUF_OUT_BOARD] = sp ? cfg->out_pos_cap * 16 : 0;
    out[OTH_BUF_OUT_PI] = sp ? cfg->out_pos_cap * OTH_NUM_ACTIONS * 4 : 0;
    out[OTH_BUF_OUT_VALUE] = sp ? cfg->out_pos_cap * 8 : 0;
    out[OTH_BUF_OUT_META] = sp ? cfg->out_pos_cap * 8 : 0;
    out[OTH_BUF_OUT_GAMES] = sp ? cfg->out_game_cap * 32 : 0;
    out[OTH_BUF_COUNTERS] = 16 * 8;
    out[OTH_BUF_SLOT_COUNTERS] = G * CNT_LOCAL * 4;
    out[OTH_BUF_HOT] = G * 256;
    out[OTH_BUF_MOVE_FLAGS] = ((G + 63) / 64) * 64;
    out[OTH_BUF_MOVE_LIST] = (4 + G) * 4;
    return OTH_OK;
}

extern "C" int oth_mcts_reset(const oth_mcts_config* cfg, const oth_mcts_buffers* b, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    int e = cuda_status(cudaMemsetAsync(p.counters, 0, 16 * 8, (cudaStream_t)stream));
    if (e != OTH_OK) return e;
    e = cuda_status(cudaMemsetAsync(p.move_list, 0, 16, (cudaStream_t)stream));
    if (e != OTH_OK) return e;
    LAUNCH_LANES(k_mcts_reset, mcts_grid(cfg), stream, p);
    return cuda_status(cudaGetLastError());
}

extern "C" int oth_mcts_set_roots_masked(const oth_mcts_config* cfg, const oth_mcts_buffers* b, const uint64_t* own, const uint64_t* opp,
                                         const int8_t* players, const uint8_t* mask, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    if (!own || !opp || !players) return OTH_E_ARG;
    LAUNCH_LANES(k_mcts_set_roots, mcts_grid(cfg), stream, p, (const u64*)own, (const u64*)opp, players, mask);
    return cuda_status(cudaGetLastError());
}

extern "C" int oth_mcts_set_roots(const oth_mcts_config* cfg, const oth_mcts_buffers* b, const uint64_t* own, const uint64_t* opp,
                                  const int8_t* players, void* stream)
{
    return oth_mcts_set_roots_masked(cfg, b, own, opp, players, nullptr, stream);
}

extern "C" int oth_mcts_begin_search_masked(const oth_mcts_config* cfg, const oth_mcts_buffers* b, const uint8_t* mask, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    k_mcts_begin_search<<<(cfg->n_slots + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, mask);
    return cuda_status(cudaGetLastError());
}

extern "C" int oth_mcts_begin_search(const oth_mcts_config* cfg, const oth_mcts_buffers* b, void* stream)
{
    return oth_mcts_begin_search_masked(cfg, b, nullptr, stream);
}

// ---- per-launch kernel timing (othello_b200_experimental.h): CUDA events recorded on the launching stream around
// the step kernel and around the move kernel.  The state lives in a handle the caller attaches to its
// oth_mcts_buffers -- one per engine, nothing global.
namespace {
struct LaunchProfile {
    cudaEvent_t* ev = nullptr;  // 3 per launch: before step, after step, after move
    int cap = 0, n = 0;
};

inline bool prof_mark(const oth_mcts_buffers* b, int which, void* stream)
{
    LaunchProfile* pr = (LaunchProfile*)b->profile;
    if (!pr || pr->n >= pr->cap) return false;
    cudaEventRecord(pr->ev[3 * pr->n + which], (cudaStream_t)stream);
    return true;
}
inline void prof_next(const oth_mcts_buffers* b) { ((LaunchProfile*)b->profile)->n++; }
}  // namespace

extern "C" int oth_mcts_profile_create(int32_t max_launches, void** out_handle)
{
    if (max_launches <= 0 || max_launches > (1 << 20) || !out_handle) return OTH_E_ARG;
    LaunchProfile* pr = new LaunchProfile;
    pr->ev = new cudaEvent_t[3 * (size_t)max_launches];
    for (int i = 0; i < 3 * max_launches; i++)
        if (cudaEventCreate(&pr->ev[i]) != cudaSuccess) {
            for (int k = 0; k < i; k++) cudaEventDestroy(pr->ev[k]);
            delete[] pr->ev;
            delete pr;
            return cuda_status(cudaGetLastError());
        }
    pr->cap = max_launches;
    *out_handle = pr;
    return OTH_OK;
}

extern "C" int oth_mcts_profile_read(void* handle, float* step_ms, float* move_ms, int32_t* n_launches)
{
    LaunchProfile* pr = (LaunchProfile*)handle;
    if (!pr || !n_launches) return OTH_E_ARG;
    const int n = pr->n;
    int rc = OTH_OK;
    for (int i = 0; i < n && rc == OTH_OK; i++) {
        if (cudaEventSynchronize(pr->ev[3 * i + 2]) != cudaSuccess) rc = cuda_status(cudaGetLastError());
        float a = 0.f, c = 0.f;
        if (rc == OTH_OK && (cudaEventElapsedTime(&a, pr->ev[3 * i], pr->ev[3 * i + 1]) != cudaSuccess ||
                             cudaEventElapsedTime(&c, pr->ev[3 * i + 1], pr->ev[3 * i + 2]) != cudaSuccess))
            rc = cuda_status(cudaGetLastError());
        if (step_ms) step_ms[i] = a;
        if (move_ms) move_ms[i] = c;
    }
    *n_launches = n;
    pr->n = 0;
    return rc;
}

extern "C" int oth_mcts_profile_destroy(void* handle)
{
    LaunchProfile* pr = (LaunchProfile*)handle;
    if (!pr) return OTH_E_ARG;
    for (int i = 0; i < 3 * pr->cap; i++) cudaEventDestroy(pr->ev[i]);
    delete[] pr->ev;
    delete pr;
    return OTH_OK;
}

// The move kernel after a step kernel: flag scan (move_launch = 0) or due list (1).
static inline void launch_move(const oth_mcts_config* cfg, const Params& p, void* stream)
{
    if (!cfg->self_play) return;
    if (cfg->move_launch) {
        const int blocks = (cfg->n_slots + kBlock / 32 - 1) / (kBlock / 32);  // a warp per slot if every slot is due
        const int cap = sm_count() * 4;
        k_mcts_move_list<<<blocks < cap ? blocks : cap, kBlock, 0, (cudaStream_t)stream>>>(p);
    } else {
        k_mcts_move<<<move_grid(cfg), kBlock, 0, (cudaStream_t)stream>>>(p);
    }
}

extern "C" int oth_mcts_step(const oth_mcts_config* cfg, const oth_mcts_buffers* b, const float* priors, const float* values,
                             float* nn_input, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    if (cfg->eval_kind == OTH_EVAL_EXTERNAL && (!priors || !values || !nn_input)) return OTH_E_ARG;
    p.priors = priors;
    p.values = values;
    p.nn_input = nn_input;
    nvtxRangePushA("oth_mcts_step");
    const bool prof = prof_mark(b, 0, stream);
    if (cfg->eval_kind != OTH_EVAL_EXTERNAL && !cfg->split_stub) {
        LAUNCH_LANES(k_mcts_step_stub, mcts_grid(cfg), stream, p);
        if (prof) prof_mark(b, 1, stream);
    } else {
        if (cfg->eval_kind != OTH_EVAL_EXTERNAL) LAUNCH_LANES(k_mcts_step_devstub, mcts_grid(cfg), stream, p);
        else LAUNCH_LANES(k_mcts_step, mcts_grid(cfg), stream, p);
        if (prof) prof_mark(b, 1, stream);
        launch_move(cfg, p, stream);
    }
    if (prof) {
        prof_mark(b, 2, stream);
        prof_next(b);
    }
    nvtxRangePop();
    return cuda_status(cudaGetLastError());
}

extern "C" int oth_mcts_step_fused(const oth_mcts_config* cfg, const oth_mcts_buffers* b, const void* logits, int64_t logits_stride,
                                   const void* value_preact, int64_t value_stride, int32_t is_bf16, float* priors_out, float* values_out,
                                   float* nn_input, void* stream)
{
    return oth_mcts_step_fused_mapped(cfg, b, logits, logits_stride, value_preact, value_stride, is_bf16, nullptr, priors_out, values_out,
                                      nn_input, stream);
}

extern "C" int oth_mcts_step_fused_mapped(const oth_mcts_config* cfg, const oth_mcts_buffers* b, const void* logits, int64_t logits_stride,
                                          const void* value_preact, int64_t value_stride, int32_t is_bf16, const int32_t* eval_map,
                                          float* priors_out, float* values_out, float* nn_input, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    if (cfg->eval_kind != OTH_EVAL_EXTERNAL || !logits || !value_preact || !nn_input || logits_stride < OTH_NUM_ACTIONS || value_stride < 1 ||
        (priors_out == nullptr) != (values_out == nullptr))
        return OTH_E_ARG;
    p.logits = logits;
    p.vpre = value_preact;
    p.logits_stride = (int)logits_stride;
    p.vpre_stride = (int)value_stride;
    p.raw_bf16 = is_bf16 ? 1 : 0;
    p.eval_map = eval_map;
    p.priors_out = priors_out;
    p.values_out = values_out;
    p.nn_input = nn_input;
    nvtxRangePushA("oth_mcts_step_fused");
    const bool prof = prof_mark(b, 0, stream);
    LAUNCH_LANES(k_mcts_step_fused, mcts_grid(cfg), stream, p);
    if (prof) prof_mark(b, 1, stream);
    launch_move(cfg, p, stream);
    if (prof) {
        prof_mark(b, 2, stream);
        prof_next(b);
    }
    nvtxRangePop();
    return cuda_status(cudaGetLastError());
}

extern "C" int oth_mcts_poll(const oth_mcts_config* cfg, const oth_mcts_buffers* b, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    // everything except the two live output-ring counters is re-derived
    int e = cuda_status(cudaMemsetAsync(p.counters, 0, OTH_CNT_POSITIONS * 8, (cudaStream_t)stream));
    if (e != OTH_OK) return e;
    e = cuda_status(cudaMemsetAsync(p.counters + OTH_CNT_MOVES, 0, (16 - OTH_CNT_MOVES) * 8, (cudaStream_t)stream));
    if (e != OTH_OK) return e;
    k_mcts_poll<<<sm_count(), 256, 0, (cudaStream_t)stream>>>(p);
    return cuda_status(cudaGetLastError());
}

extern "C" int oth_mcts_advance(const oth_mcts_config* cfg, const oth_mcts_buffers* b, const int32_t* actions, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    if (!actions) return OTH_E_ARG;
    LAUNCH_LANES(k_mcts_advance, mcts_grid(cfg), stream, p, actions);
    return cuda_status(cudaGetLastError());
}

extern "C" int oth_mcts_root_stats(const oth_mcts_config* cfg, const oth_mcts_buffers* b, int32_t* counts, double* child_value,
                                   double* child_prior, double* root_value, int32_t* root_n, uint64_t* root_board, void* stream)
{
    Params p;
    const int rc = make_params(cfg, b, &p);
    if (rc != OTH_OK) return rc;
    LAUNCH_LANES(k_mcts_root_stats, mcts_grid(cfg), stream, p, counts, child_value, child_prior, root_value, root_n, (u64*)root_board);
    return cuda_status(cudaGetLastError());
}
