// Evaluation de-duplication for the batched leaf evaluation (engine side of the network boundary).
//
// The reference evaluates one leaf per simulation per game (MCTS_model.py:325-336).  Thousands of concurrent
// games that all start from the same position ask, for the first plies, for the very same positions: measured on
// whole C4 games 97 % of the leaves of a batch are duplicates of one another in plies 0-4, 67 % in plies 4-8, none
// after ply 16 (profiles/r02_dedup_probe_c4.json).  This pass finds the distinct pending leaves of a batch,
// writes their canonical planes to a compact [bucket][64] network input and gives every slot the row of its
// position (eval_map), so the network runs on `bucket` rows instead of n_slots and the step kernel reads
// logits[eval_map[slot]].  A slot's own sequence of evaluations is unchanged -- same positions, same order -- so the
// search results depend on it only through the network's outputs (the parity contract: "given identical network
// outputs"); slots whose position did not fit the bucket simply wait one more launch.
//
//   k_dedup_keys   32-bit hash of the pending leaf (own, opp) per slot; non-waiting slots sort to the end
//   cub radix sort (hash, slot)
//   k_dedup_heads  first slot of every run of IDENTICAL boards (full 128-bit compare; a hash collision only splits a
//                  group, it never merges different positions)
//   cub inclusive scan -> row index of every group
//   k_dedup_emit   eval_map[slot], representative slot per row, counters
//   k_dedup_planes canonical planes of the representatives (Models.py:16: +1 own, -1 opp)
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/othello_b200.h"
#include "common.cuh"

using namespace oth;

namespace {

typedef unsigned long long u64;
constexpr uint32_t kNotWaiting = 0xffffffffu;

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

struct DedupWs {  // carved out of the caller's workspace
    uint32_t *keys_in, *keys_out, *slots_in, *slots_out;
    int32_t *heads, *incl, *rep;
    uint32_t* launch_counter;
    void* cub;
    size_t cub_bytes;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t cub_temp_bytes(int n)
{
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, n);
    return a > b ? a : b;
}

size_t carve(void* base, int n, DedupWs* w)
{
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align256(bytes);
        return (void*)q;
    };
    const size_t cb = cub_temp_bytes(n);
    void* a0 = take((size_t)n * 4);
    void* a1 = take((size_t)n * 4);
    void* a2 = take((size_t)n * 4);
    void* a3 = take((size_t)n * 4);
    void* a4 = take((size_t)n * 4);
    void* a5 = take((size_t)n * 4);
    void* a6 = take((size_t)n * 4);
    void* a7 = take(256);
    void* a8 = take(cb);
    if (w) {
        w->keys_in = (uint32_t*)a0;
        w->keys_out = (uint32_t*)a1;
        w->slots_in = (uint32_t*)a2;
        w->slots_out = (uint32_t*)a3;
        w->heads = (int32_t*)a4;
        w->incl = (int32_t*)a5;
        w->rep = (int32_t*)a6;
        w->launch_counter = (uint32_t*)a7;
        w->cub = a8;
        w->cub_bytes = cb;
    }
    return off;
}

// hot record of slot s: {leaf_own, leaf_opp, ...} in its first 16 bytes (mcts_kernels.cu SlotHot)
__device__ __forceinline__ ulonglong2 leaf_board(const uint4* hot, uint32_t s)
{
    const uint4 h = hot[(size_t)s * 16];
    return make_ulonglong2(((u64)h.y << 32) | h.x, ((u64)h.w << 32) | h.z);
}

__global__ void __launch_bounds__(256) k_dedup_keys(const oth_mcts_ctl* __restrict__ ctl, const uint4* __restrict__ hot, int n,
                                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ slots,
                                                    const uint32_t* __restrict__ launch_counter)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t key = kNotWaiting;
    if (ctl[s].phase == OTH_PH_WAIT_EVAL && ctl[s].top >= 1) {
        const ulonglong2 b = leaf_board(hot, (uint32_t)s);
        // salted per launch: which positions miss a too-small bucket changes from launch to launch (no starvation)
        key = (uint32_t)(mix64(b.x ^ mix64(b.y ^ (u64)*launch_counter)) >> 32);
        if (key == kNotWaiting) key = kNotWaiting - 1;
    }
    keys[s] = key;
    slots[s] = (uint32_t)s;
}

__global__ void __launch_bounds__(256) k_dedup_heads(const uint4* __restrict__ hot, int n, const uint32_t* __restrict__ keys,
                                                     const uint32_t* __restrict__ slots, int32_t* __restrict__ heads)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    int head = 0;
    if (k != kNotWaiting) {
        head = 1;
        if (i > 0 && keys[i - 1] == k) {
            const ulonglong2 a = leaf_board(hot, slots[i]), b = leaf_board(hot, slots[i - 1]);
            head = !(a.x == b.x && a.y == b.y);
        }
    }
    heads[i] = head;
}

__global__ void __launch_bounds__(256) k_dedup_emit(int n, int bucket, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ slots,
                                                    const int32_t* __restrict__ heads, const int32_t* __restrict__ incl,
                                                    int32_t* __restrict__ eval_map, int32_t* __restrict__ rep, int32_t* __restrict__ stats,
                                                    uint32_t* __restrict__ launch_counter)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slots[i];
    int row = -1;
    if (keys[i] != kNotWaiting) {
        const int u = incl[i] - 1;
        if (u < bucket) {
            row = u;
            if (heads[i]) rep[u] = (int32_t)s;
        }
    }
    if (eval_map) eval_map[s] = row;
    if (i == n - 1) {
        stats[0] = incl[i];  // distinct pending leaves in this batch
        *launch_counter += 1;
    }
    if (keys[i] != kNotWaiting && (i == n - 1 || keys[i + 1] == kNotWaiting)) stats[1] = i + 1;  // slots waiting for an evaluation
    if (i == 0 && keys[0] == kNotWaiting) stats[1] = 0;
}

__global__ void __launch_bounds__(256) k_dedup_planes(const uint4* __restrict__ hot, int bucket, const int32_t* __restrict__ rep,
                                                      const int32_t* __restrict__ stats, float* __restrict__ compact_input)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int u = t >> 6, e = t & 63;
    if (u >= bucket) return;
    float v = 0.0f;
    if (u < stats[0]) {
        const ulonglong2 b = leaf_board(hot, (uint32_t)rep[u]);
        v = ((b.x >> e) & 1) ? 1.0f : (((b.y >> e) & 1) ? -1.0f : 0.0f);
    }
    compact_input[(size_t)u * 64 + e] = v;  // rows past the last distinct position are zero boards (their outputs are unused)
}

}  // namespace

extern "C" int oth_mcts_dedup_workspace_bytes(int32_t n_slots, int64_t* bytes)
{
    if (n_slots <= 0 || !bytes) return OTH_E_ARG;
    *bytes = (int64_t)carve(nullptr, n_slots, nullptr);
    return OTH_OK;
}

extern "C" int oth_mcts_dedup(const oth_mcts_config* cfg, const oth_mcts_buffers* b, int32_t bucket, void* workspace, int64_t workspace_bytes,
                              float* compact_input, int32_t* eval_map, int32_t* stats, void* stream)
{
    if (!cfg || !b || cfg->n_slots <= 0 || !workspace || !stats || bucket < 0 || bucket > cfg->n_slots) return OTH_E_ARG;
    if (bucket > 0 && (!compact_input || !eval_map)) return OTH_E_ARG;
    const int n = cfg->n_slots;
    DedupWs w;
    if ((int64_t)carve(workspace, n, &w) > workspace_bytes) return OTH_E_ARG;
    const oth_mcts_ctl* ctl = (const oth_mcts_ctl*)b->buf[OTH_BUF_CTL];
    const uint4* hot = (const uint4*)b->buf[OTH_BUF_HOT];
    if (!ctl || !hot) return OTH_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = (n + 255) / 256;
    k_dedup_keys<<<g, 256, 0, st>>>(ctl, hot, n, w.keys_in, w.slots_in, w.launch_counter);
    size_t cb = w.cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub, cb, w.keys_in, w.keys_out, w.slots_in, w.slots_out, n, 0, 32, st);
    if (e != cudaSuccess) return cuda_status(e);
    k_dedup_heads<<<g, 256, 0, st>>>(hot, n, w.keys_out, w.slots_out, w.heads);
    cb = w.cub_bytes;
    e = cub::DeviceScan::InclusiveSum(w.cub, cb, w.heads, w.incl, n, st);
    if (e != cudaSuccess) return cuda_status(e);
    k_dedup_emit<<<g, 256, 0, st>>>(n, bucket, w.keys_out, w.slots_out, w.heads, w.incl, bucket > 0 ? eval_map : nullptr, w.rep, stats,
                                    w.launch_counter);
    if (bucket > 0) k_dedup_planes<<<(bucket * 64 + 255) / 256, 256, 0, st>>>(hot, bucket, w.rep, stats, compact_input);
    return cuda_status(cudaGetLastError());
}
