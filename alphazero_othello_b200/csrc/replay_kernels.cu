// Replay-buffer ingest on the GPU: duplicate aggregation of replay tuples.
//
// Restates Trainer._aggregate_duplicates (train.py:142-173): tuples with the same
// (state, model_version) are collapsed into one sample whose policy is the arithmetic mean of
// the duplicates' policies, re-normalised, and whose value is the mean value; buckets come out in
// order of first occurrence.  The reference keys buckets by sha1(int8 board) -- here the packed
// canonical board (own, opp) IS the identity, no hashing.
//
// Bit-exactness with the reference's sequential numpy loop:
//   sum_pi += pi    float32, in buffer order     -> stable LSD radix sort keeps buffer order
//   sum_v  += v     float64 (Python float)          inside each bucket; one warp walks a bucket
//   avg_pi = sum_pi / count ; avg_pi /= (avg_pi.sum() + 1e-12)   float32, numpy pairwise sum
//   v = float32(sum_v / count)
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/othello_b200.h"
#include "common.cuh"

using namespace oth;

namespace {

typedef unsigned long long u64;

__global__ void __launch_bounds__(256) k_rp_init(const int32_t* __restrict__ versions, u64* __restrict__ key, unsigned* __restrict__ idx,
                                                 int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        key[i] = (u64)(unsigned)versions[i];
        idx[i] = (unsigned)i;
    }
}

__global__ void __launch_bounds__(256) k_rp_gather_key(const u64* __restrict__ boards, const unsigned* __restrict__ idx, int word,
                                                       u64* __restrict__ key, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = boards[2 * (int64_t)idx[i] + word];
}

__global__ void __launch_bounds__(256) k_rp_heads(const u64* __restrict__ boards, const int32_t* __restrict__ versions,
                                                  const unsigned* __restrict__ idx, int* __restrict__ flag, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int f = 1;
    if (k > 0) {
        const int64_t a = idx[k], b = idx[k - 1];
        f = !(boards[2 * a] == boards[2 * b] && boards[2 * a + 1] == boards[2 * b + 1] && versions[a] == versions[b]);
    }
    flag[k] = f;
}

// seg[k] = inclusive scan of flags; bucket s starts where flag is set
__global__ void __launch_bounds__(256) k_rp_segments(const int* __restrict__ flag, const int* __restrict__ seg, const unsigned* __restrict__ idx,
                                                     int* __restrict__ seg_start, u64* __restrict__ first_key, unsigned* __restrict__ seg_ids,
                                                     long long* __restrict__ out_m, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int m = seg[n - 1];
    if (k == 0) *out_m = m;
    if (flag[k]) {
        const int s = seg[k] - 1;
        seg_start[s] = (int)k;
        first_key[s] = idx[k];  // stable sort: the head of a bucket is its earliest tuple
        seg_ids[s] = (unsigned)s;
    }
    if (k >= m) {  // pad so a full-length sort pushes unused entries to the end
        first_key[k] = ~0ULL;
        seg_ids[k] = (unsigned)k;
    }
    if (k == n - 1) seg_start[m] = (int)n;
}

// One warp per output bucket (rank r in first-occurrence order).
__global__ void __launch_bounds__(256) k_rp_accumulate(const u64* __restrict__ boards, const float* __restrict__ pis,
                                                       const double* __restrict__ values, const int32_t* __restrict__ versions,
                                                       const unsigned* __restrict__ idx, const int* __restrict__ seg_start,
                                                       const unsigned* __restrict__ order, const long long* __restrict__ m_ptr,
                                                       u64* __restrict__ out_boards, float* __restrict__ out_pis, float* __restrict__ out_values,
                                                       int32_t* __restrict__ out_versions, int32_t* __restrict__ out_counts)
{
    __shared__ float sh[8][OTH_NUM_ACTIONS + 3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long m = *m_ptr;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < m; r += nw) {
        const int s = (int)order[r];
        const int k0 = seg_start[s], k1 = seg_start[s + 1];
        const int cnt = k1 - k0;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;  // actions lane, lane+32, 64 (lane 0)
        double sv = 0.0;
        for (int k = k0; k < k1; k++) {
            const int64_t i = idx[k];
            const float* p = pis + i * OTH_NUM_ACTIONS;
            if (k == k0) {  // sum_pi = pi.copy(); sum_v = v
                a0 = p[lane];
                a1 = p[lane + 32];
                if (lane == 0) {
                    a2 = p[64];
                    sv = values[i];
                }
            } else {
                a0 = __fadd_rn(a0, p[lane]);
                a1 = __fadd_rn(a1, p[lane + 32]);
                if (lane == 0) {
                    a2 = __fadd_rn(a2, p[64]);
                    sv = __dadd_rn(sv, values[i]);
                }
            }
        }
        const float fc = (float)cnt;
        __syncwarp();
        sh[w][lane] = __fdiv_rn(a0, fc);
        sh[w][lane + 32] = __fdiv_rn(a1, fc);
        if (lane == 0) sh[w][64] = __fdiv_rn(a2, fc);
        __syncwarp();
        // numpy pairwise float32 sum over 65 entries (8 strided accumulators + fixed tree + tail)
        float t = 0.0f;
        if (lane < 8) {
            t = sh[w][lane];
#pragma unroll
            for (int q = 1; q < 8; q++) t = __fadd_rn(t, sh[w][8 * q + lane]);
        }
        t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
        t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
        t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 4));
        t = __shfl_sync(0xffffffffu, t, 0);
        t = __fadd_rn(t, sh[w][64]);
        const float den = __fadd_rn(t, (float)1e-12);
        float* op = out_pis + r * OTH_NUM_ACTIONS;
        op[lane] = __fdiv_rn(sh[w][lane], den);
        op[lane + 32] = __fdiv_rn(sh[w][lane + 32], den);
        if (lane == 0) {
            op[64] = __fdiv_rn(sh[w][64], den);
            const int64_t i0 = idx[k0];
            out_boards[2 * r] = boards[2 * i0];
            out_boards[2 * r + 1] = boards[2 * i0 + 1];
            out_values[r] = (float)__ddiv_rn(sv, (double)cnt);
            out_versions[r] = versions[i0];
            out_counts[r] = cnt;
        }
        __syncwarp();
    }
}

struct Layout {
    size_t key_a, key_b, idx_a, idx_b, flag, seg, seg_start, first_a, first_b, ord_a, ord_b, cub, total;
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int make_layout(int64_t n, Layout* L, size_t* cub_bytes)
{
    size_t tmp1 = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp1, (u64*)nullptr, (u64*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, tmp2, (int*)nullptr, (int*)nullptr, (int)n);
    *cub_bytes = tmp1 > tmp2 ? tmp1 : tmp2;
    size_t o = 0;
    const size_t N = (size_t)n + 1;
    L->key_a = o; o += align_up(N * 8);
    L->key_b = o; o += align_up(N * 8);
    L->idx_a = o; o += align_up(N * 4);
    L->idx_b = o; o += align_up(N * 4);
    L->flag = o; o += align_up(N * 4);
    L->seg = o; o += align_up(N * 4);
    L->seg_start = o; o += align_up(N * 4);
    L->first_a = o; o += align_up(N * 8);
    L->first_b = o; o += align_up(N * 8);
    L->ord_a = o; o += align_up(N * 4);
    L->ord_b = o; o += align_up(N * 4);
    L->cub = o; o += align_up(*cub_bytes);
    L->total = o;
    return OTH_OK;
}

}  // namespace

extern "C" int oth_replay_aggregate_workspace_bytes(int64_t n, int64_t* bytes)
{
    if (n < 0 || n > 0x7fffffff || !bytes) return OTH_E_ARG;
    Layout L;
    size_t cb;
    make_layout(n < 1 ? 1 : n, &L, &cb);
    *bytes = (int64_t)L.total;
    return OTH_OK;
}

extern "C" int oth_replay_aggregate(const uint64_t* boards, const float* pis, const double* values, const int32_t* versions, int64_t n,
                                    void* workspace, int64_t workspace_bytes, uint64_t* out_boards, float* out_pis, float* out_values,
                                    int32_t* out_versions, int32_t* out_counts, int64_t* out_m, void* stream_)
{
    if (n < 0 || n > 0x7fffffff || !out_m) return OTH_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    if (n == 0) return cuda_status(cudaMemsetAsync(out_m, 0, 8, st));
    if (!boards || !pis || !values || !versions || !workspace || !out_boards || !out_pis || !out_values || !out_versions || !out_counts)
        return OTH_E_ARG;
    Layout L;
    size_t cb;
    make_layout(n, &L, &cb);
    if ((size_t)workspace_bytes < L.total) return OTH_E_ARG;
    char* w = (char*)workspace;
    u64 *key_a = (u64*)(w + L.key_a), *key_b = (u64*)(w + L.key_b);
    unsigned *idx_a = (unsigned*)(w + L.idx_a), *idx_b = (unsigned*)(w + L.idx_b);
    int *flag = (int*)(w + L.flag), *seg = (int*)(w + L.seg), *seg_start = (int*)(w + L.seg_start);
    u64 *first_a = (u64*)(w + L.first_a), *first_b = (u64*)(w + L.first_b);
    unsigned *ord_a = (unsigned*)(w + L.ord_a), *ord_b = (unsigned*)(w + L.ord_b);
    void* tmp = w + L.cub;
    const int g = (int)((n + 255) / 256);
    const u64* bd = (const u64*)boards;
#define CUB_OK(x)                                    \
    do {                                             \
        cudaError_t _e = (x);                        \
        if (_e != cudaSuccess) return cuda_status(_e); \
    } while (0)
    // stable LSD sort by (own, opp, version): version first, own last
    k_rp_init<<<g, 256, 0, st>>>(versions, key_a, idx_a, n);
    CUB_OK(cub::DeviceRadixSort::SortPairs(tmp, cb, key_a, key_b, idx_a, idx_b, (int)n, 0, 32, st));
    k_rp_gather_key<<<g, 256, 0, st>>>(bd, idx_b, 1, key_a, n);
    CUB_OK(cub::DeviceRadixSort::SortPairs(tmp, cb, key_a, key_b, idx_b, idx_a, (int)n, 0, 64, st));
    k_rp_gather_key<<<g, 256, 0, st>>>(bd, idx_a, 0, key_a, n);
    CUB_OK(cub::DeviceRadixSort::SortPairs(tmp, cb, key_a, key_b, idx_a, idx_b, (int)n, 0, 64, st));
    // idx_b: tuples grouped by key, buffer order inside a group
    k_rp_heads<<<g, 256, 0, st>>>(bd, versions, idx_b, flag, n);
    CUB_OK(cub::DeviceScan::InclusiveSum(tmp, cb, flag, seg, (int)n, st));
    k_rp_segments<<<g, 256, 0, st>>>(flag, seg, idx_b, seg_start, first_a, ord_a, (long long*)out_m, n);
    // buckets in order of first occurrence (dict insertion order)
    CUB_OK(cub::DeviceRadixSort::SortPairs(tmp, cb, first_a, first_b, ord_a, ord_b, (int)n, 0, 64, st));
    k_rp_accumulate<<<grid_for(n * 32 < (int64_t)sm_count() * 2048 ? n * 32 : (int64_t)sm_count() * 2048, 256), 256, 0, st>>>(
        bd, pis, values, versions, idx_b, seg_start, ord_b, (const long long*)out_m, (u64*)out_boards, out_pis, out_values, out_versions,
        out_counts);
#undef CUB_OK
    return cuda_status(cudaGetLastError());
}
