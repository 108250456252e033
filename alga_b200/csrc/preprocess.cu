// ReadPreprocess::getPrefixReads on the GPU (reference: src/IO/ReadPreprocess.cpp:13-152, called from main.cpp:132-134):
// which reads are duplicates of, or prefixes of, another read -- the step right before the overlap-graph build.
//
// The reference sorts all reads (bit strings LSB first, then length, then id) and marks read i when its successor
// starts with it.  A read r is a prefix of its successor iff it is a prefix of ANY read s that sorts after it, i.e.
//     prefix(s, len r) == r   and   (len s, id s) > (len r, id r),
// because the reads that start with r form one contiguous run that begins with the copies of r itself.  That
// condition needs no order, only a dictionary: all reads go into a hash table keyed by (content, length); every read s
// rolls a hash over its own prefixes and looks up, at every length that occurs in the read set, who equals that
// prefix.  Per read r the lookups leave two facts:
//     DUP     an identical read with a greater id exists      (its successor is that copy: same length)
//     LONGER  a longer read starts with r
// and getPrefixReads follows: type 2 removes r if DUP or LONGER, and also its reverse complement (id ^ 1,
// Read.cpp:229-236) if LONGER and not DUP (then r's successor IS longer, ReadPreprocess.cpp:43-47); type 1 removes r
// if DUP.  Exact word compares decide, the hash only finds candidates.
#include <cstdarg>
#include <cstdio>

#include "../../include/alga_gpu.h"
#include "launch.h"

namespace alga {

namespace {

inline int grid_for(uint64_t n_items, int per_block, const LaunchCfg &cfg, int max_blocks_per_sm = 16) {
    uint64_t need = (n_items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t) cfg.sm_count * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

constexpr uint32_t kMaxLen = 65535;             // lengths are tracked in a 64 Kbit map
constexpr uint32_t kLenWords = (kMaxLen + 1) / 32;
constexpr uint64_t kBase = 0x9E3779B97F4A7C15ull;  // odd: code_j * kBase^j mod 2^64
constexpr uint32_t kDup = 1u, kLonger = 2u;

__device__ __forceinline__ uint64_t key_of(uint64_t h, uint32_t len) { return mix64(h ^ ((uint64_t) len * 0xD6E8FEB86659FD93ull)); }

__global__ void length_map_kernel(ReadsDev R, uint32_t *lenmap, uint32_t *too_long) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < R.n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t len = R.len[i];
        if (len == 0) continue;
        if (len > kMaxLen) {
            *too_long = 1u;
            continue;
        }
        if (!((lenmap[len >> 5] >> (len & 31u)) & 1u)) atomicOr(lenmap + (len >> 5), 1u << (len & 31u));
    }
}

// *uniform = the one length that occurs in the read set, or 0 if several do
__global__ void length_info_kernel(const uint32_t *__restrict__ lenmap, uint32_t *uniform) {
    __shared__ uint32_t cnt, which;
    if (threadIdx.x == 0) cnt = 0, which = 0;
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < kLenWords; k += blockDim.x) {
        const uint32_t w = lenmap[k];
        if (w) {
            atomicAdd(&cnt, (uint32_t) __popc(w));
            atomicMax(&which, 32u * k + (uint32_t) (__ffs(w) - 1));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *uniform = cnt == 1 ? which : 0u;
}

// All reads have one length (the usual case: equal-length reads after trimming): only duplicates can occur, one lookup per
// read, and the key needs no rolling structure -- a hash of the blocks, one multiply per 16 nucleotides.
__device__ __forceinline__ uint64_t block_hash(const uint32_t *__restrict__ p, uint32_t len) {
    uint64_t h = 0x243F6A8885A308D3ull;
    const uint32_t nw = (len + 15u) >> 4;
    for (uint32_t k = 0; k < nw; k++) h = (h ^ __ldg(p + k)) * kBase + (h >> 29);
    return h;
}

__global__ void insert_reads_kernel(ReadsDev R, SeedTable t, const uint32_t *__restrict__ uniform) {
    const uint32_t ulen = *uniform;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < R.n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t len = R.len[i];
        if (len == 0 || len > kMaxLen) continue;
        const uint32_t *p = read_ptr(R, (uint32_t) i);
        if (ulen) {
            insert_seed(t, key_of(block_hash(p, len), len), (uint32_t) i);
            continue;
        }
        uint64_t h = 0, pw = 1;
        for (uint32_t j0 = 0; j0 < len; j0 += 16) {  // one block (16 nucleotides) per load
            uint32_t w = __ldg(p + (j0 >> 4));
            const uint32_t nj = len - j0 < 16u ? len - j0 : 16u;
            for (uint32_t j = 0; j < nj; j++, w >>= 2) {
                h += (uint64_t) (w & 3u) * pw;
                pw *= kBase;
            }
        }
        insert_seed(t, key_of(h, len), (uint32_t) i);
    }
}

// first `len` nucleotides of s == read r (of length len)?
__device__ __forceinline__ bool same_prefix(const uint32_t *__restrict__ ps, const uint32_t *__restrict__ pr, uint32_t len) {
    const uint32_t nbits = 2u * len, nw = (nbits + 31u) >> 5;
    for (uint32_t k = 0; k < nw; k++) {
        uint32_t x = __ldg(ps + k) ^ __ldg(pr + k);
        if (k == nw - 1 && (nbits & 31u)) x &= (1u << (nbits & 31u)) - 1u;
        if (x) return false;
    }
    return true;
}

__global__ void probe_prefixes_kernel(ReadsDev R, SeedTable t, const uint32_t *__restrict__ lenmap, uint32_t *flags,
                                      const uint32_t *__restrict__ uniform) {
    const uint32_t ulen = *uniform;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < R.n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t s = (uint32_t) i, len_s = R.len[s];
        if (len_s == 0 || len_s > kMaxLen) continue;
        const uint32_t *ps = read_ptr(R, s);
        if (ulen) {  // every copy of s with a smaller id is a duplicate (the greatest id of a set of equal reads stays)
            probe_seed(t, key_of(block_hash(ps, len_s), len_s), [&](uint32_t r) {
                if (r >= s) return;
                if (!same_prefix(ps, read_ptr(R, r), len_s)) return;
                atomicOr(flags + r, kDup);
            });
            continue;
        }
        uint64_t h = 0, pw = 1;
        uint32_t lm = __ldg(lenmap);  // the word of the length map that holds bit l (reloaded every 32 lengths)
        for (uint32_t j0 = 0; j0 < len_s; j0 += 16) {  // one block (16 nucleotides) per load
            uint32_t w = __ldg(ps + (j0 >> 4));
            const uint32_t nj = len_s - j0 < 16u ? len_s - j0 : 16u;
            for (uint32_t j = 0; j < nj; j++, w >>= 2) {
                h += (uint64_t) (w & 3u) * pw;
                pw *= kBase;
                const uint32_t l = j0 + j + 1;
                if ((l & 31u) == 0) lm = __ldg(lenmap + (l >> 5));
                if (!((lm >> (l & 31u)) & 1u)) continue;  // no read has this length
                probe_seed(t, key_of(h, l), [&](uint32_t r) {
                    if (r == s || R.len[r] != l) return;
                    if (l == len_s && r > s) return;  // r sorts after s: s does not remove it
                    if (!same_prefix(ps, read_ptr(R, r), l)) return;
                    atomicOr(flags + r, l < len_s ? kLonger : kDup);
                });
            }
        }
    }
}

__global__ void mark_removed_kernel(const uint32_t *__restrict__ flags, uint32_t n, int remove_type, uint8_t *mask) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t f = flags[i];
        if (remove_type == 1) {
            if (f & kDup) mask[i] = 1;
        } else if (f) {
            mask[i] = 1;
            if ((f & kLonger) && !(f & kDup) && (i ^ 1ull) < n) mask[i ^ 1ull] = 1;  // several writers, one value
        }
    }
}

}  // namespace

// flags: n words (zeroed here), lenmap: prefix_reads_lenmap_words() words (zeroed here; the last word reports a read that is too long)
void launch_prefix_reads(const ReadsDev &R, const SeedTable &t, int remove_type, uint32_t *lenmap, uint32_t *flags,
                         uint8_t *mask, cudaStream_t s, const LaunchCfg &cfg) {
    cudaMemsetAsync(lenmap, 0, (size_t) (kLenWords + 2) * 4, s);
    cudaMemsetAsync(flags, 0, (size_t) (R.n ? R.n : 1) * 4, s);
    cudaMemsetAsync(mask, 0, R.n ? R.n : 1, s);
    if (!R.n) return;
    length_map_kernel<<<grid_for(R.n, 256, cfg), 256, 0, s>>>(R, lenmap, lenmap + kLenWords + 1);
    length_info_kernel<<<1, 256, 0, s>>>(lenmap, lenmap + kLenWords);
    insert_reads_kernel<<<grid_for(R.n, 128, cfg), 128, 0, s>>>(R, t, lenmap + kLenWords);
    probe_prefixes_kernel<<<grid_for(R.n, 128, cfg), 128, 0, s>>>(R, t, lenmap, flags, lenmap + kLenWords);
    mark_removed_kernel<<<grid_for(R.n, 256, cfg), 256, 0, s>>>(flags, R.n, remove_type, mask);
    if (cfg.launches) *cfg.launches += 5;
}
// [0, kLenWords) the map, [kLenWords] the uniform length (or 0), [kLenWords + 1] "a read is too long"
size_t prefix_reads_lenmap_words() { return kLenWords + 2; }

}  // namespace alga
