// 2-bit packing, reference fingerprints and candidate verification kernels (sm_100a).
#include "launch.h"

namespace alga {

namespace {


inline int grid_for(uint64_t n_items, int per_block, const LaunchCfg &cfg, int max_blocks_per_sm = 16) {
    uint64_t need = (n_items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t) cfg.sm_count * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}
inline void bump(const LaunchCfg &cfg) {
    if (cfg.launches) (*cfg.launches)++;
}

// ------------------------------------------------------------------------------------------------
// Read::createSequence (Read.cpp:40-68): 'C' -> 01, 'G' -> 10, 'T' -> 11, anything else -> 00.
// The ASCII stream of a tile of reads is staged into shared memory with aligned 128-bit loads, then
// every thread assembles one 32-bit block (16 nucleotides).
constexpr int kPackThreads = 256;
constexpr int kPackTileBytes = 32 * 1024;

__device__ __forceinline__ uint32_t nt_code(uint8_t ch) {
    return ch == 'C' ? 1u : (ch == 'G' ? 2u : (ch == 'T' ? 3u : 0u));
}

__global__ void __launch_bounds__(kPackThreads)
pack_reads_kernel(const uint8_t *__restrict__ ascii, uint32_t n_reads, uint32_t len_nt, uint32_t reads_per_tile,
                  uint32_t *__restrict__ words) {
    __shared__ uint4 stage[kPackTileBytes / 16 + 2];
    const uint32_t wpr = (len_nt + 15) / 16;
    const uint64_t n_tiles = (n_reads + reads_per_tile - 1) / reads_per_tile;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t r0 = tile * reads_per_tile;
        const uint64_t r1 = min((uint64_t) n_reads, r0 + reads_per_tile);
        const uint64_t byte0 = r0 * len_nt, byte1 = r1 * len_nt;
        const uint64_t a0 = byte0 & ~15ull;  // aligned start (the buffer itself is 256-byte aligned)
        const uint32_t n_vec = (uint32_t) ((byte1 - a0 + 15) / 16);
        const uint4 *src = reinterpret_cast<const uint4 *>(ascii + a0);
        for (uint32_t v = threadIdx.x; v < n_vec; v += blockDim.x) stage[v] = __ldg(src + v);
        __syncthreads();
        const uint8_t *sb = reinterpret_cast<const uint8_t *>(stage) + (byte0 - a0);
        const uint32_t n_words = (uint32_t) (r1 - r0) * wpr;
        for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
            const uint32_t r = w / wpr, k = w - r * wpr;
            const uint32_t first = k * 16;
            const uint32_t cnt = min(16u, len_nt - first);
            const uint8_t *p = sb + (uint64_t) r * len_nt + first;
            uint32_t word = 0;
            for (uint32_t j = 0; j < cnt; j++) word |= nt_code(p[j]) << (2 * j);
            words[(r0 + r) * wpr + k] = word;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Reference fingerprints (GraphCreatorPrefSuf.cpp:213-236): h = sum s_j 4^j mod M over the length-L
// prefix / suffix, M64 = 10^18+3, M32 = 10^9+7.  One thread per read; same recurrences as the reference.
constexpr uint64_t kM64 = 1000000000000000003ull;
constexpr uint32_t kM32 = 1000000007u;

__device__ __forceinline__ uint32_t nt_at(const uint32_t *__restrict__ p, uint32_t j) {
    return (__ldg(p + (j >> 4)) >> ((j & 15u) * 2u)) & 3u;
}
__device__ __forceinline__ uint64_t red64(uint64_t x) {  // x < 4*M64 + 4
    if (x >= 2 * kM64) x -= 2 * kM64;
    if (x >= kM64) x -= kM64;
    return x;
}

__global__ void fingerprints_kernel(ReadsDev R, int L, uint64_t *__restrict__ pre64, uint32_t *__restrict__ pre32,
                                    uint64_t *__restrict__ suf64, uint32_t *__restrict__ suf32) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < R.n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t len = R.len[i];
        if (L < 1 || (uint32_t) L > len) continue;
        const uint32_t *p = read_ptr(R, (uint32_t) i);
        uint64_t h64 = 0, f64 = 1, g64 = 0;
        uint64_t h32 = 0, f32 = 1, g32 = 0;
        for (int l = 1; l <= L; l++) {
            const uint32_t c = nt_at(p, (uint32_t) l - 1);       // updatePrefixHash
            h64 = red64(h64 + c * f64);
            h32 = (h32 + c * f32) % kM32;
            f64 = red64(f64 * 4);
            f32 = (f32 * 4) % kM32;
            const uint32_t e = nt_at(p, len - (uint32_t) l);     // updateSuffixHash
            g64 = red64(g64 * 4 + e);
            g32 = (g32 * 4 + e) % kM32;
        }
        pre64[i] = h64;
        pre32[i] = (uint32_t) h32;
        suf64[i] = g64;
        suf32[i] = (uint32_t) g32;
    }
}

// ------------------------------------------------------------------------------------------------
// AlignmentControllerHybrid::canAlign -> AlignmentControllerLowErrorRate::canAlign, one pair per warp.
// x = (a >> 2*off) XOR b over the first 2*ov bits; lanes take 32-bit words of x round-robin, popcounts
// are combined with warp shuffles.
__device__ __forceinline__ uint32_t range_mask(int64_t word, int64_t lo_bit, int64_t hi_bit /*inclusive*/) {
    const int64_t w0 = word * 32, w1 = w0 + 31;
    if (hi_bit < w0 || lo_bit > w1 || hi_bit < lo_bit) return 0u;
    const int lo = (int) (lo_bit > w0 ? lo_bit - w0 : 0);
    const int hi = (int) (hi_bit < w1 ? hi_bit - w0 : 31);
    const uint32_t upto_hi = hi == 31 ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u);
    return upto_hi & ~((1u << lo) - 1u);
}

__global__ void __launch_bounds__(256)
verify_pairs_kernel(ReadsDev R, const int32_t *__restrict__ pairs, uint64_t n_pairs, VerifyDev V,
                    uint8_t *__restrict__ verdict) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < n_pairs; i += n_warps) {
        const int32_t a = pairs[3 * i], b = pairs[3 * i + 1], off = pairs[3 * i + 2];
        bool ok = a >= 0 && b >= 0 && (uint32_t) a < R.n && (uint32_t) b < R.n && off >= 0;
        int64_t la = 0, lb = 0, ov = 0;
        if (ok) {
            la = R.len[a];
            lb = R.len[b];
            ok = la > 0 && lb > 0;
        }
        if (ok) {
            ov = (la < lb + off ? la : lb + off) - off;
            ok = 100ll * off <= (int64_t) V.max_offset_pct * la && off >= V.min_offset && ov >= V.min_overlap_area &&
                 lb + off - la >= 0 && ov > 0;
        }
        uint32_t diff = 0, ends = 0;
        if (ok) {
            const uint32_t *pa = read_ptr(R, (uint32_t) a);
            const uint32_t *pb = read_ptr(R, (uint32_t) b);
            const int64_t nbits = 2 * ov;
            const int64_t n_words = (nbits + 31) >> 5;
            for (int64_t k = lane; k < n_words; k += 32) {
                uint32_t x = bits32(pa, (uint32_t) (2 * off + 32 * k)) ^ __ldg(pb + k);
                x &= range_mask(k, 0, nbits - 1);
                diff += __popc(x);
                ends += __popc(x & range_mask(k, 0, 2ll * V.same_ends));                   // count(0, SAME<<1)
                ends += __popc(x & range_mask(k, 2 * (ov - V.same_ends), nbits - 1));      // last SAME nucleotides
            }
        }
        for (int d = 16; d; d >>= 1) {
            diff += __shfl_xor_sync(kFull, diff, d);
            ends += __shfl_xor_sync(kFull, ends, d);
        }
        if (lane == 0) {
            uint8_t v = 0;
            if (ok && ends == 0) {
                const int64_t sim = (2 * ov - (int64_t) diff) >> 1;
                v = 100 * sim >= (int64_t) V.threshold_pct * ov ? 1 : 0;
            }
            if (ok && !v && V.lcs_rate_pct > 0) v = 2;  // passed the filters, failed the low-error test: on to the banded LCS
            verdict[i] = v;
        }
    }
}

// AlignmentControllerLCS::canAlign (AlignmentControllerLCS.cpp:30-59) for the pairs marked 2 above: longest common subsequence
// inside the band |q - (p - offset)| <= E (calculateLCS :61-150), accepted iff 100 * lcs > rate * overlap.  A row of the
// band depends on the row before and, cell by cell, on its left neighbour, so one pair is one sequential walk: one THREAD per
// pair, the two rows of 2E + 3 cells in registers (cells the reference never wrote read as 0 there: its table is a hash map).
constexpr int kLcsMaxBand = 8;
__device__ __forceinline__ uint32_t nt_of(const uint32_t *__restrict__ p, int64_t pos) {
    return (__ldg(p + (pos >> 4)) >> (2 * (pos & 15))) & 3u;
}
__global__ void __launch_bounds__(128)
lcs_pairs_kernel(ReadsDev R, const int32_t *__restrict__ pairs, uint64_t n_pairs, VerifyDev V, uint8_t *__restrict__ verdict) {
    const int E = V.lcs_band;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t) gridDim.x * blockDim.x) {
        if (verdict[i] != 2) continue;
        const int32_t a = pairs[3 * i], b = pairs[3 * i + 1], off = pairs[3 * i + 2];
        const int64_t la = R.len[a], lb = R.len[b];
        const int64_t ov = (la < lb + off ? la : lb + off) - off;
        const uint32_t *pa = read_ptr(R, (uint32_t) a), *pb = read_ptr(R, (uint32_t) b);
        int32_t prev[2 * kLcsMaxBand + 3], cur[2 * kLcsMaxBand + 3];  // index d + E + 1, d = q - (p - off)
#pragma unroll
        for (int k = 0; k < 2 * kLcsMaxBand + 3; k++) prev[k] = cur[k] = 0;
        const int64_t p_beg = off - E > 0 ? off - E : 0;
        const int64_t p_last = la - 1 < lb - 1 + off ? la - 1 : lb - 1 + off;
        for (int64_t pp = p_beg; pp <= p_last; pp++) {
            const uint32_t ca = nt_of(pa, pp);
#pragma unroll
            for (int k = 0; k < 2 * kLcsMaxBand + 3; k++) {
                const int d = k - E - 1;
                int32_t v = 0;
                if (k >= 1 && k <= 2 * E + 1) {
                    const int64_t q = pp - off + d;
                    if (q >= 0 && q <= lb - 1) {
                        if (ca == nt_of(pb, q)) {
                            v = (pp > 0 && q > 0) ? prev[k] + 1 : 1;
                        } else {
                            if (pp > 0) v = max(v, prev[k + 1 < 2 * kLcsMaxBand + 3 ? k + 1 : k]);
                            if (q > 0) v = max(v, cur[k - 1]);
                        }
                    }
                }
                cur[k] = v;
            }
#pragma unroll
            for (int k = 0; k < 2 * kLcsMaxBand + 3; k++) prev[k] = cur[k];
        }
        int64_t lcs = 0;
        if (p_last >= p_beg) {
            const int64_t q_last = lb - 1 < p_last - off + E ? lb - 1 : p_last - off + E;
            const int64_t d = q_last - (p_last - off);
            if (d >= -E && d <= E) {
#pragma unroll
                for (int k = 0; k < 2 * kLcsMaxBand + 3; k++)
                    if (k == (int) d + E + 1) lcs = prev[k];
            }
        }
        verdict[i] = 100 * lcs > (int64_t) V.lcs_rate_pct * ov ? 1 : 0;
    }
}

// Aligned copy of a fixed-stride read set: read i -> out[i * stride_out .. + words), zero up to the stride.  One thread
// per 16 bytes of output (stride_out is a multiple of 8 words, `out` 16-byte aligned): four 4-byte loads -- consecutive
// threads, consecutive addresses up to the stride change -- and one 16-byte store.
__global__ void repack_reads_kernel(const uint32_t *__restrict__ in, uint32_t stride_in, uint32_t words, uint64_t n_reads,
                                    uint32_t *__restrict__ out, uint32_t stride_out) {
    const uint32_t q_per_read = stride_out >> 2;
    const uint64_t total = n_reads * q_per_read;
    for (uint64_t j = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; j < total; j += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t i = j / q_per_read;
        const uint32_t w = (uint32_t) (j - i * q_per_read) << 2;
        const uint32_t *p = in + i * stride_in + w;
        uint4 v;
        v.x = w < words ? __ldg(p) : 0u;
        v.y = w + 1 < words ? __ldg(p + 1) : 0u;
        v.z = w + 2 < words ? __ldg(p + 2) : 0u;
        v.w = w + 3 < words ? __ldg(p + 3) : 0u;
        reinterpret_cast<uint4 *>(out)[j] = v;
    }
}

}  // namespace

void launch_repack_reads(const uint32_t *in, uint32_t stride_in, uint32_t words, uint64_t n_reads, uint32_t *out,
                         uint32_t stride_out, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_reads) return;
    repack_reads_kernel<<<grid_for(n_reads * (stride_out >> 2), 256, cfg, 8), 256, 0, s>>>(in, stride_in, words, n_reads, out, stride_out);
    bump(cfg);
}

void launch_pack_reads(const uint8_t *ascii, uint32_t n_reads, uint32_t len_nt, uint32_t *words, cudaStream_t s,
                       const LaunchCfg &cfg) {
    if (!n_reads || !len_nt) return;
    uint32_t rpt = (kPackTileBytes - 16) / len_nt;
    if (rpt < 1) rpt = 1;  // caller guarantees len_nt <= kPackTileBytes - 16
    const uint64_t n_tiles = ((uint64_t) n_reads + rpt - 1) / rpt;
    pack_reads_kernel<<<grid_for(n_tiles, 1, cfg, 4), kPackThreads, 0, s>>>(ascii, n_reads, len_nt, rpt, words);
    bump(cfg);
}

void launch_fingerprints(const ReadsDev &R, int L, uint64_t *pre64, uint32_t *pre32, uint64_t *suf64, uint32_t *suf32,
                         cudaStream_t s, const LaunchCfg &cfg) {
    if (!R.n) return;
    fingerprints_kernel<<<grid_for(R.n, 128, cfg), 128, 0, s>>>(R, L, pre64, pre32, suf64, suf32);
    bump(cfg);
}

void launch_verify_pairs(const ReadsDev &R, const int32_t *pairs, uint64_t n_pairs, const VerifyDev &V,
                         uint8_t *verdict, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_pairs) return;
    verify_pairs_kernel<<<grid_for(n_pairs, 8, cfg, 8), 256, 0, s>>>(R, pairs, n_pairs, V, verdict);
    bump(cfg);
    if (V.lcs_rate_pct > 0) {
        lcs_pairs_kernel<<<grid_for(n_pairs, 128, cfg, 8), 128, 0, s>>>(R, pairs, n_pairs, V, verdict);
        bump(cfg);
    }
}

}  // namespace alga
