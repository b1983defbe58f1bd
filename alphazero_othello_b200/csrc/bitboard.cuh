// Othello bitboard primitives for sm_100a device code.
//
// Square numbering is the Game API's action index: bit i = row*8 + col
// (row 0 = top, col 0 = left), so "ascending action order" is ascending bit
// order.  The reference's internal numbering (envs/othello.py:336-356,
// bit = 63 - idx) is the bit reversal of this one; only the hash stub needs it.
//
// Rules restated from envs/othello.py:147-200 (_BitBoard._shift/_legal_moves/
// make_move) as parallel-prefix fills instead of the reference's 7-stage
// per-direction flood: a run of opponent discs is at most 6 long on an 8x8
// board, which three doubling steps cover.
#pragma once
#include <stdint.h>

namespace oth {

typedef unsigned long long u64;

constexpr u64 NOT_EDGE_COLS = 0x7E7E7E7E7E7E7E7EULL;  // cols 1..6: runs that may continue horizontally/diagonally
constexpr u64 INIT_BLACK = (1ULL << 28) | (1ULL << 35);  // state[3,4] = state[4,3] = +1 (envs/othello.py:136-140 via :374-392)
constexpr u64 INIT_WHITE = (1ULL << 27) | (1ULL << 36);  // state[3,3] = state[4,4] = -1

// Legal-move set for `own` (side to move) against `opp`: envs/othello.py:157-166.
__device__ __forceinline__ u64 legal_moves(u64 own, u64 opp)
{
    u64 moves = 0;
    const u64 mh = opp & NOT_EDGE_COLS;
#define OTH_AXIS(m, d)                                                   \
    {                                                                    \
        u64 l = (m) & (own << (d)), r = (m) & (own >> (d));              \
        l |= (m) & (l << (d));                                           \
        r |= (m) & (r >> (d));                                           \
        const u64 pl = (m) & ((m) << (d)), pr = (m) & ((m) >> (d));      \
        l |= pl & (l << (2 * (d)));                                      \
        r |= pr & (r >> (2 * (d)));                                      \
        l |= pl & (l << (2 * (d)));                                      \
        r |= pr & (r >> (2 * (d)));                                      \
        moves |= (l << (d)) | (r >> (d));                                \
    }
    OTH_AXIS(mh, 1)
    OTH_AXIS(opp, 8)
    OTH_AXIS(mh, 7)
    OTH_AXIS(mh, 9)
#undef OTH_AXIS
    return moves & ~(own | opp);
}

// Discs captured when `own` plays on the empty square bit `x` (one bit set):
// envs/othello.py:182-193.  Returns 0 if nothing is bracketed (illegal move).
__device__ __forceinline__ u64 flips(u64 own, u64 opp, u64 x)
{
    u64 captured = 0;
    const u64 mh = opp & NOT_EDGE_COLS;
#define OTH_RAY(m, d)                                                    \
    {                                                                    \
        u64 l = (m) & (x << (d)), r = (m) & (x >> (d));                  \
        l |= (m) & (l << (d));                                           \
        r |= (m) & (r >> (d));                                           \
        const u64 pl = (m) & ((m) << (d)), pr = (m) & ((m) >> (d));      \
        l |= pl & (l << (2 * (d)));                                      \
        r |= pr & (r >> (2 * (d)));                                      \
        l |= pl & (l << (2 * (d)));                                      \
        r |= pr & (r >> (2 * (d)));                                      \
        captured |= ((l << (d)) & own) ? l : 0ULL;                       \
        captured |= ((r >> (d)) & own) ? r : 0ULL;                       \
    }
    OTH_RAY(mh, 1)
    OTH_RAY(opp, 8)
    OTH_RAY(mh, 7)
    OTH_RAY(mh, 9)
#undef OTH_RAY
    return captured;
}

// Index of the n-th (0-based) set bit of m, ascending; m must have > n bits set.
__device__ __forceinline__ int nth_set_bit(u64 m, int n)
{
    unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
    int base = 0;
    unsigned w = lo;
    int c = __popc(lo);
    if (n >= c) {
        n -= c;
        w = hi;
        base = 32;
    }
    // w has > n bits: narrow by halves
    int p = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const unsigned low = (w >> p) & ((1u << s) - 1u);
        const int cl = __popc(low);
        if (n >= cl) {
            n -= cl;
            p += s;
        }
    }
    return base + p;
}

// Result of moving: boards from the NEXT side-to-move's perspective, as
// _BitBoard.make_move leaves them (envs/othello.py:195-200; pass :177-180).
struct Board {
    u64 own, opp;
};

__device__ __forceinline__ Board apply_move(u64 own, u64 opp, int action, u64 f)
{
    Board b;
    if (action == 64) {
        b.own = opp;
        b.opp = own;
    } else {
        b.own = opp ^ f;
        b.opp = own | (1ULL << action) | f;
    }
    return b;
}

// get_value_and_terminated (envs/othello.py:435-454) for side-to-move `own`
// given its already computed legal set: returns 0 = continue, 1 = must pass,
// 2 = terminal; *value = sign(#own - #opp) when terminal.
__device__ __forceinline__ int position_status(u64 own, u64 opp, u64 own_moves, int* value)
{
    *value = 0;
    if (own_moves) return 0;
    if (legal_moves(opp, own)) return 1;
    const int d = __popcll(own) - __popcll(opp);
    *value = (d > 0) - (d < 0);
    return 2;
}

}  // namespace oth
