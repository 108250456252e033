// InputReader::readInput on the GPU (reference: src/IO/InputReader.cpp:44-139, 142-180, 272-391) and the renumbering of
// the surviving reads (main.cpp:150-232): the two steps that turn the input files into the read set the graph creators
// see.  SURVEY.md 8-f rank 2.
//
// The file text is copied to HBM as it is.  Records are found without any sequential pass:
//   marks      every thread looks at 16 bytes and counts line ends ('\n'; plain input: token starts), a scan over the
//              per-block counts gives every mark its global number k, and mark k with k % lines_per_record == 0 is the
//              line end in front of the sequence line of record k / lines_per_record (readOneRead1, :142-180);
//   scan       one warp per record: end of the line, the space stripping and end trimming of readParallelJob (:286-303),
//              the character check (:316-334), N detection (:341-347) and the minimal-period filter (:343-353) as
//              "is there a period p <= 20", 32 positions per step;  reading stops at the first record whose sequence line
//              is empty (:284), so the smallest such record number is kept with atomicMin;
//   pack       one warp per record, one 16-nucleotide word per lane: forward strand and reverse complement
//              (Read::createSequence, Read.cpp:40-68; getComplimentaryString, InputReader.cpp:23-33) written straight
//              to their final ids -- reverse complement at the even id, mates of the two files interleaved (:54-85).
// All results are defined by the reference's --threads=1 order (with several threads a malformed file is read
// differently, see oracle/input_oracle.c).
#include <cstdio>
#include <cstdlib>

#include "../../include/alga_gpu.h"
#include "launch.h"

namespace alga {

namespace {

constexpr int kMarkThreads = 256;
constexpr int kMarkChunk = kMarkThreads * 16;  // bytes of text per block

__device__ __forceinline__ bool is_space_c(uint32_t c) { return c == 32u || (c - 9u) <= 4u; }  // isspace in the C locale

// marks of the 16 bytes at `base` (bit j = byte base + j): line ends, or (plain input, `str >> s`) token starts
template <bool PLAIN>
__device__ __forceinline__ uint32_t marks16(const uint8_t *__restrict__ text, uint64_t n, uint64_t base) {
    if (base >= n) return 0;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + base));  // the buffer is padded to a multiple of 16
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
    bool prev_space = true;
    if (PLAIN && base > 0) prev_space = is_space_c(__ldg(text + base - 1));
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t c = (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
        if (base + j < n) {
            if (PLAIN) {
                const bool sp = is_space_c(c);
                if (!sp && prev_space) m |= 1u << j;
                prev_space = sp;
            } else if (c == '\n') {
                m |= 1u << j;
            }
        }
    }
    return m;
}

// exclusive prefix of `v` over the block (kMarkThreads threads); *total = block sum
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_sum[kMarkThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sum[wid] = inc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int i = 0; i < kMarkThreads / 32; i++) {
        const uint32_t s = warp_sum[i];
        if (i < wid) before += s;
        all += s;
    }
    *total = all;
    return before + inc - v;
}

template <bool PLAIN>
__global__ void __launch_bounds__(kMarkThreads) count_marks_kernel(const uint8_t *__restrict__ text, uint64_t n,
                                                                   uint32_t *__restrict__ block_cnt) {
    const uint64_t base = (uint64_t) blockIdx.x * kMarkChunk + (uint64_t) threadIdx.x * 16u;
    const uint32_t c = (uint32_t) __popc(marks16<PLAIN>(text, n, base));
    uint32_t total;
    block_exclusive(c, &total);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = total;
}

// rec_start[r] = first byte of the sequence of record r; rec_end[r] = the line end behind it (mark lpr * r + 1, if the
// file has one; not written for plain input)
template <bool PLAIN>
__global__ void __launch_bounds__(kMarkThreads) write_marks_kernel(const uint8_t *__restrict__ text, uint64_t n,
                                                                   const uint64_t *__restrict__ block_off, uint32_t lpr,
                                                                   uint64_t *__restrict__ rec_start,
                                                                   uint64_t *__restrict__ rec_end, uint64_t n_rec) {
    const uint64_t base = (uint64_t) blockIdx.x * kMarkChunk + (uint64_t) threadIdx.x * 16u;
    uint32_t m = marks16<PLAIN>(text, n, base);
    uint32_t total;
    uint64_t k = block_off[blockIdx.x] + block_exclusive((uint32_t) __popc(m), &total);
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        if (PLAIN) {
            if (k < n_rec) rec_start[k] = base + j;
        } else {
            const uint64_t r = k / lpr, q = k - r * lpr;
            if (q == 0 && r < n_rec) rec_start[r] = base + j + 1;
            if (q == 1 && r < n_rec) rec_end[r] = base + j;
        }
        k++;
    }
}

struct __align__(16) RecInfo {
    uint64_t begin;   // first byte of the trimmed sequence
    uint32_t len;     // its length
    uint32_t status;  // kRec* bits, 0 = a read
};
constexpr uint32_t kRecHasN = 1u, kRecStr = 2u, kRecEmpty = 4u, kRecBad = 8u;
constexpr uint32_t kRecNull = 15u;  // any of the above: no read
constexpr uint32_t kRecSlow = 16u;  // a read, but not one the word-parallel path handles (U, stripped spaces, > 512 nt)

struct InputScalars {
    uint32_t first_empty;  // smallest record number whose sequence line is empty (reading stops there)
    uint32_t first_bad;    // smallest record number with a character other than A C G T N U
    uint32_t max_len;
    uint32_t pad;
    unsigned long long n_with_n, n_str;
    unsigned long long sum_len, n_alive;  // over the records that became reads (Global::calculateAvgReadLength, Global.h:133-145)
};

// first position in [from, to) whose byte satisfies pred, or `to`; all 32 lanes call it with the same arguments
template <class P>
__device__ __forceinline__ uint64_t warp_find(const uint8_t *__restrict__ text, uint64_t from, uint64_t to, int lane, P pred) {
    for (uint64_t p = from; p < to; p += 32) {
        const uint64_t q = p + lane;
        const bool hit = q < to && pred((uint32_t) __ldg(text + q));
        const unsigned m = __ballot_sync(kFull, hit);
        if (m) return p + (uint64_t) (__ffs(m) - 1);
    }
    return to;
}

__device__ __forceinline__ uint32_t sym_of(uint32_t c, int rna) { return (rna && c == 'U') ? (uint32_t) 'T' : c; }

// ---- word-parallel helpers: 16 characters per lane ------------------------------------------------------------------
// 16 bytes at an arbitrary address as four little-endian words (reads up to 7 bytes past them: the text buffer is padded)
__device__ __forceinline__ void load16(const uint8_t *__restrict__ p, uint32_t (&c)[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t) 3);
    const uint32_t sh = (uint32_t) (a & 3u) * 8u;
    uint32_t v[5];
#pragma unroll
    for (int j = 0; j < 5; j++) v[j] = __ldg(q + j);
#pragma unroll
    for (int j = 0; j < 4; j++) c[j] = __funnelshift_r(v[j], v[j + 1], sh);
}
// four characters -> four 2-bit codes (A 0, C 1, G 2, T 3) in the low byte; *bad gets a non-zero byte for every
// character that is not A, C, G or T.  (c >> 1) & 3 maps A C T G to 0 1 2 3, "ACTG"[that] must give the character back.
__device__ __forceinline__ uint32_t codes4(uint32_t c, uint32_t *bad) {
    const uint32_t x = (c >> 1) & 0x03030303u;
    uint32_t t = (x | (x >> 4)) & 0x00FF00FFu;
    const uint32_t sel = (t | (t >> 8)) & 0xFFFFu;
    *bad = __byte_perm(0x47544341u, 0u, sel) ^ c;
    const uint32_t k = x ^ ((x >> 1) & 0x01010101u);
    return (k | (k >> 6) | (k >> 12) | (k >> 18)) & 0xFFu;
}
// 16 characters -> one packed block; bad != 0 iff one of the first `cnt` characters is not A, C, G or T
__device__ __forceinline__ uint32_t pack16(const uint32_t (&c)[4], uint32_t cnt, uint32_t *bad) {
    uint32_t w = 0, b = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
        uint32_t bj;
        w |= codes4(c[j], &bj) << (8u * j);
        const uint32_t nv = cnt > 4u * j ? cnt - 4u * j : 0u;
        b |= nv >= 4u ? bj : (bj & ((1u << (8u * nv)) - 1u));
    }
    *bad = b;
    return cnt >= 16u ? w : (w & ((1u << (2u * cnt)) - 1u));
}
// order of the sixteen 2-bit fields reversed
__device__ __forceinline__ uint32_t reverse_fields(uint32_t p) {
    const uint32_t r = __brev(p);
    return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}
constexpr uint32_t kFastMaxLen = 512;  // one block per lane

// The general form of one record (any length, spaces, N, U): sequential warp loops over the bytes.
template <bool PLAIN>
__device__ __forceinline__ void scan_record_generic(const uint8_t *__restrict__ text, uint64_t n, uint64_t start, uint32_t r,
                                                    int trim_left, int trim_right, int rna, int str_threshold, int lane,
                                                    RecInfo *__restrict__ info, InputScalars *__restrict__ sc) {
    RecInfo out;
    out.begin = start, out.len = 0, out.status = kRecEmpty;
    // the sequence line / token (readOneRead1)
    uint64_t end = PLAIN ? warp_find(text, start, n, lane, [](uint32_t c) { return is_space_c(c); })
                         : warp_find(text, start, n, lane, [](uint32_t c) { return c == '\n'; });
    if (end == start) {  // s == "": the reader stops here (InputReader.cpp:284)
        if (lane == 0) {
            atomicMin(&sc->first_empty, r);
            info[r] = out;
        }
        return;
    }
    // :286-291 -- leading spaces off, cut at the next space
    uint64_t b = start, e = end;
    if (!PLAIN) {
        b = warp_find(text, start, end, lane, [](uint32_t c) { return c != ' '; });
        e = warp_find(text, b, end, lane, [](uint32_t c) { return c == ' '; });
    }
    // :298-303 -- end trimming unless the read is short
    if (e - b >= (uint64_t) (trim_left + trim_right + 10)) {
        b += (uint64_t) trim_left;
        e -= (uint64_t) trim_right;
    }
    const uint64_t len64 = e - b;
    // :316-334 -- characters
    bool bad = false, has_n = false;
    for (uint64_t i = lane; i < len64; i += 32) {
        const uint32_t c = __ldg(text + b + i);
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N' && c != 'U') bad = true;
        if (c == 'N') has_n = true;
    }
    bad = __any_sync(kFull, bad) || len64 > 0x7FFFFFFFull;
    has_n = __any_sync(kFull, has_n);
    uint32_t status = kRecSlow;
    if (bad) {
        status = kRecBad;
        if (lane == 0) atomicMin(&sc->first_bad, r);
    } else if (has_n) {
        status = kRecHasN;
    } else if (len64 == 0) {
        status = kRecStr;  // an all-space line: MinPeriod("") is undefined in the reference; dropped
    } else {
        // :343-353 -- MinPeriod(s) <= threshold  <=>  some p <= min(threshold, len) has s[i] == s[i + p] for all i
        const uint32_t len = (uint32_t) len64;
        const uint32_t pmax = (uint32_t) str_threshold < len ? (uint32_t) str_threshold : len;
        for (uint32_t p = 1; p <= pmax && status == kRecSlow; p++) {
            bool periodic = true;
            for (uint32_t i0 = 0; i0 + p < len; i0 += 32) {
                const uint32_t i = i0 + lane;
                bool ok = true;
                if (i + p < len) ok = sym_of(__ldg(text + b + i), rna) == sym_of(__ldg(text + b + i + p), rna);
                if (!__all_sync(kFull, ok)) {
                    periodic = false;
                    break;
                }
            }
            if (periodic) status = kRecStr;
        }
    }
    if (lane == 0) {
        out.begin = b, out.len = (uint32_t) len64, out.status = status;
        info[r] = out;
    }
}

// One warp per record.  FASTA / FASTQ records of up to 512 characters made of A, C, G, T only (practically all of
// them) never loop over bytes: the line end is known from the marks, lane w loads characters 16w .. 16w+15 of the
// trimmed sequence in one go, checks and packs them with word arithmetic, and the period test "s[i] == s[i+p] for all
// i" becomes "(X >> 2p) xor X has no bit below 2 (len - p)" on the packed blocks X, one block per lane.
template <bool PLAIN>
__global__ void __launch_bounds__(256) scan_records_kernel(const uint8_t *__restrict__ text, uint64_t n,
                                                           const uint64_t *__restrict__ rec_start,
                                                           const uint64_t *__restrict__ rec_end, uint64_t n_marks, uint32_t lpr,
                                                           uint32_t n_cand, int trim_left, int trim_right, int rna,
                                                           int str_threshold, RecInfo *__restrict__ info,
                                                           InputScalars *__restrict__ sc) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    // the bounds of the next record are fetched while the current one is worked on
    uint64_t start = 0, end = n;
    if (warp0 < n_cand) {
        start = rec_start[warp0];
        if (!PLAIN && (uint64_t) lpr * warp0 + 1 < n_marks) end = rec_end[warp0];
    }
    for (uint32_t r = warp0; r < n_cand;) {
        const uint32_t r_next = r + n_warps;
        uint64_t start_next = 0, end_next = n;
        if (r_next < n_cand) {
            start_next = rec_start[r_next];
            if (!PLAIN && (uint64_t) lpr * r_next + 1 < n_marks) end_next = rec_end[r_next];
        }
        bool fast = !PLAIN && trim_left <= 32 && trim_right <= 32 && str_threshold <= 31;  // 2p < 64: two neighbour blocks suffice
        uint64_t b = start;
        uint32_t len = 0, x = 0;
        if (fast) {
            const uint64_t raw = end - start;
            fast = raw > 0 && raw <= kFastMaxLen;
            if (fast) {
                const bool trimmed = raw >= (uint64_t) (trim_left + trim_right + 10);
                b = start + (trimmed ? (uint64_t) trim_left : 0);
                len = (uint32_t) raw - (trimmed ? (uint32_t) (trim_left + trim_right) : 0u);
                bool odd = false;  // a space in the trimmed-away ends moves the token (InputReader.cpp:286-291): general form
                if (trimmed) {
                    if (lane < trim_left) odd |= __ldg(text + start + lane) == ' ';
                    if (lane < trim_right) odd |= __ldg(text + end - 1 - lane) == ' ';
                }
                const uint32_t cnt = len > 16u * lane ? min(16u, len - 16u * lane) : 0u;
                if (cnt) {
                    uint32_t c[4], bad;
                    load16(text + b + 16u * lane, c);
                    x = pack16(c, cnt, &bad);
                    odd |= bad != 0;
                }
                fast = !__any_sync(kFull, odd);
            }
        }
        if (!fast) {
            scan_record_generic<PLAIN>(text, n, start, r, trim_left, trim_right, rna, str_threshold, lane, info, sc);
        } else {
            // MinPeriod(s) <= threshold on the packed blocks (lane w holds X_w = nucleotides 16w .. 16w+15).
            // Filter: lane p tests period p on the first 64 nucleotides only (blocks 0..3, broadcast); a period of the
            // read is a period of its prefix, and a random read refutes all of them there.
            const uint32_t pmax = (uint32_t) str_threshold < len ? (uint32_t) str_threshold : len;
            const uint32_t blk[6] = {__shfl_sync(kFull, x, 0), __shfl_sync(kFull, x, 1), __shfl_sync(kFull, x, 2),
                                     __shfl_sync(kFull, x, 3), 0u, 0u};
            const uint32_t lp = len < 64u ? len : 64u;
            bool cand = false;
            if (lane >= 1 && (uint32_t) lane <= pmax) {
                const uint32_t p = (uint32_t) lane, sh = (2u * p) & 31u;
                const bool big = 2u * p >= 32u;
                uint32_t diff = 0;
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t y = __funnelshift_r(big ? blk[w + 1] : blk[w], big ? blk[w + 2] : blk[w + 1], sh);
                    const int nb = (int) (2u * (lp - p)) - 32 * w;
                    const uint32_t m = nb >= 32 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
                    diff |= (y ^ blk[w]) & m;
                }
                cand = diff == 0;
            }
            unsigned cands = __ballot_sync(kFull, cand);  // bit p: period p survives the prefix
            uint32_t status = 0;
            if (cands) {  // the whole read, one surviving period at a time
                uint32_t x1 = __shfl_down_sync(kFull, x, 1), x2 = __shfl_down_sync(kFull, x, 2);
                if (lane >= 31) x1 = 0;
                if (lane >= 30) x2 = 0;
                while (cands && !status) {
                    const uint32_t p = (uint32_t) __ffs(cands) - 1u;
                    cands &= cands - 1u;
                    const uint32_t sh = (2u * p) & 31u;
                    const uint32_t y = __funnelshift_r(2u * p >= 32u ? x1 : x, 2u * p >= 32u ? x2 : x1, sh);
                    const int nb = (int) (2u * (len - p)) - 32 * lane;  // bits of this block that take part
                    const uint32_t m = nb >= 32 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
                    if (!__any_sync(kFull, ((y ^ x) & m) != 0u)) status = kRecStr;
                }
            }
            if (lane == 0) {
                RecInfo out;
                out.begin = b, out.len = len, out.status = status;
                info[r] = out;
            }
        }
        r = r_next, start = start_next, end = end_next;
    }
}

// totals over the records that are actually read (r < n_rec)
__global__ void record_totals_kernel(const RecInfo *__restrict__ info, uint32_t n_rec, InputScalars *__restrict__ sc) {
    uint32_t mx = 0, cn = 0, cs = 0, ca = 0;
    unsigned long long sum = 0;
    for (uint64_t r = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; r < n_rec; r += (uint64_t) gridDim.x * blockDim.x) {
        const RecInfo x = info[r];
        if ((x.status & kRecNull) == 0) {
            mx = x.len > mx ? x.len : mx;
            sum += x.len;
            ca++;
        }
        cn += (x.status & kRecHasN) ? 1u : 0u;
        cs += (x.status & kRecStr) ? 1u : 0u;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(kFull, mx, d);
        mx = o > mx ? o : mx;
        cn += __shfl_xor_sync(kFull, cn, d);
        cs += __shfl_xor_sync(kFull, cs, d);
        ca += __shfl_xor_sync(kFull, ca, d);
        sum += __shfl_xor_sync(kFull, sum, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mx) atomicMax(&sc->max_len, mx);
        if (cn) atomicAdd(&sc->n_with_n, (unsigned long long) cn);
        if (cs) atomicAdd(&sc->n_str, (unsigned long long) cs);
        if (ca) {
            atomicAdd(&sc->n_alive, (unsigned long long) ca);
            atomicAdd(&sc->sum_len, sum);
        }
    }
}

// Read::createSequence codes; the reverse strand goes through getComplimentaryString first (a U stays a U there, and
// packs as 0 like every other non-ACGT character)
__device__ __forceinline__ uint32_t code_fw(uint32_t c, int rna) {
    return c == 'C' ? 1u : (c == 'G' ? 2u : ((c == 'T' || (rna && c == 'U')) ? 3u : 0u));
}
__device__ __forceinline__ uint32_t code_rc(uint32_t c, int rna) {
    return c == 'A' ? 3u : (c == 'C' ? 2u : (c == 'G' ? 1u : 0u));
}

// record r of file `file` -> ids (id0, id0 + 1) = (reverse complement, forward), id0 = r * id_step + 2 * file.
// One warp per record, one block per lane: the forward block from the 16 characters at 16w, the reverse-complement
// block from the 16 characters that end at len - 16w (fields reversed, codes complemented); records the word-parallel
// path does not take (kRecSlow, or a stride of more than 32 blocks) loop over their characters.
__global__ void __launch_bounds__(256) pack_records_kernel(const uint8_t *__restrict__ text, const RecInfo *__restrict__ info,
                                                           uint32_t n_rec, int rna, uint32_t id_step, uint32_t id_first,
                                                           uint32_t stride, uint32_t *__restrict__ words,
                                                           uint32_t *__restrict__ len_out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp0; r < n_rec; r += n_warps) {
        const RecInfo x = info[r];
        const uint64_t id_rc = (uint64_t) r * id_step + id_first, id_fw = id_rc + 1;
        const uint32_t len = (x.status & kRecNull) ? 0u : x.len;
        if (lane == 0) {
            len_out[id_rc] = len;
            len_out[id_fw] = len;
        }
        const uint8_t *s = text + x.begin;
        if (!(x.status & kRecSlow) && stride <= 32u) {
            if ((uint32_t) lane < stride) {
                uint32_t fw = 0, rc = 0;
                const uint32_t i0 = 16u * lane;
                if (i0 < len) {
                    const uint32_t cnt = min(16u, len - i0);
                    uint32_t c[4], bad;
                    load16(s + i0, c);
                    fw = pack16(c, cnt, &bad);
                    // reverse-complement positions i0 .. i0+15 are forward positions len-1-i0 down to len-16-i0
                    const int f0 = (int) len - (int) i0 - 16;
                    load16(s + (f0 > 0 ? f0 : 0), c);
                    uint32_t p = pack16(c, 16u, &bad);
                    if (f0 < 0) p <<= 2 * (-f0);
                    rc = ~reverse_fields(p);
                    if (cnt < 16u) rc &= (1u << (2u * cnt)) - 1u;
                }
                words[id_fw * stride + lane] = fw;
                words[id_rc * stride + lane] = rc;
            }
            continue;
        }
        for (uint32_t w = lane; w < stride; w += 32) {
            uint32_t fw = 0, rc = 0;
            const uint32_t i0 = w * 16u;
            if (i0 < len) {
#pragma unroll
                for (uint32_t j = 0; j < 16; j++) {
                    const uint32_t i = i0 + j;
                    if (i < len) {
                        fw |= code_fw(__ldg(s + i), rna) << (2u * j);
                        rc |= code_rc(__ldg(s + (len - 1u - i)), rna) << (2u * j);
                    }
                }
            }
            words[id_fw * stride + w] = fw;
            words[id_rc * stride + w] = rc;
        }
    }
}

// ---- renumbering (main.cpp:150-232) ---------------------------------------------------------------------------
// A unit = the two strands (2u, 2u + 1) of one record.  Unit u survives iff READS[2u] != nullptr (:168); a surviving
// unit whose second read is gone is the reference's assert (:173) and reported through *err.
struct RemapScalars {
    uint32_t max_len;
    uint32_t err;  // 1 + id of a read that survives without its reverse complement
};

__device__ __forceinline__ bool alive(const ReadsDev &R, const uint8_t *__restrict__ mask, uint64_t i) {
    return R.len[i] != 0 && !(mask && mask[i]);
}

__global__ void remap_flags_kernel(ReadsDev R, const uint8_t *__restrict__ mask, uint32_t n_units, uint32_t *__restrict__ flag,
                                   RemapScalars *__restrict__ sc) {
    uint32_t mx = 0;
    for (uint64_t u = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; u < n_units; u += (uint64_t) gridDim.x * blockDim.x) {
        const bool a = alive(R, mask, 2 * u);
        flag[u] = a ? 1u : 0u;
        if (a) {
            if (!alive(R, mask, 2 * u + 1)) atomicMax(&sc->err, (uint32_t) (2 * u) + 1u);
            const uint32_t l0 = R.len[2 * u], l1 = R.len[2 * u + 1];
            mx = max(mx, max(l0, l1));
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(kFull, mx, d));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(&sc->max_len, mx);
}

__global__ void remap_scatter_kernel(ReadsDev R, uint32_t n_units, const uint32_t *__restrict__ flag,
                                     const uint32_t *__restrict__ pos, uint32_t stride, uint32_t *__restrict__ words,
                                     uint32_t *__restrict__ len_out, uint32_t *__restrict__ old_id,
                                     uint8_t *__restrict__ paired_offset, uint32_t min_keep_len) {
    const uint64_t per_unit = 2ull * stride, total = (uint64_t) n_units * per_unit;
    for (uint64_t idx = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; idx < total; idx += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t u = (uint32_t) (idx / per_unit);
        if (!flag[u]) continue;
        const uint32_t rem = (uint32_t) (idx - (uint64_t) u * per_unit), k = rem / stride, w = rem - k * stride;
        const uint32_t src = 2u * u + k;
        const uint64_t dst = 2ull * pos[u] + k;
        const uint32_t len = R.len[src];
        words[dst * stride + w] = (w < (len + 15u) / 16u && len >= min_keep_len) ? read_ptr(R, src)[w] : 0u;
        if (w == 0) {
            len_out[dst] = len < min_keep_len ? 0u : len;  // main.cpp:253-266: too short for the graph creators -> nullptr
            old_id[dst] = src;
            // Global::pairedReadOffset (:176-197): 1 / 2 for the first / second mate when both survive, else 0
            uint8_t po = 0;
            if ((u & 1u) == 0) po = (u + 1 < n_units && flag[u + 1]) ? 1 : 0;
            else po = flag[u - 1] ? 2 : 0;
            paired_offset[dst] = po;
        }
    }
}

// Device -> page-locked host memory with ordinary stores (the host buffer is mapped into the device's address space).
// Used instead of the DMA engine for the bulk download that overlaps the graph build: the small control read-backs of
// the build go through that engine and would otherwise queue behind the bulk transfer.
__global__ void __launch_bounds__(256) copy_to_host_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint64_t bytes) {
    const uint64_t n16 = bytes / 16;
    const uint4 *s16 = reinterpret_cast<const uint4 *>(src);
    uint4 *d16 = reinterpret_cast<uint4 *>(dst);
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n16; i += (uint64_t) gridDim.x * blockDim.x) d16[i] = __ldg(s16 + i);
    if (blockIdx.x == 0 && threadIdx.x < (bytes & 15u)) dst[n16 * 16 + threadIdx.x] = src[n16 * 16 + threadIdx.x];
}

inline int grid_for(uint64_t n_items, int per_block, const LaunchCfg &cfg, int max_blocks_per_sm = 8) {
    uint64_t need = (n_items + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t) cfg.sm_count * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

}  // namespace

uint64_t input_mark_blocks(uint64_t n_bytes) { return (n_bytes + kMarkChunk - 1) / kMarkChunk; }
size_t input_rec_info_bytes() { return sizeof(RecInfo); }
size_t input_scalars_bytes() { return sizeof(InputScalars); }

// block_cnt[input_mark_blocks(n)] = marks per 4 KiB block of text (text padded to a multiple of 16 bytes)
void launch_count_marks(const uint8_t *text, uint64_t n, bool plain, uint32_t *block_cnt, cudaStream_t s, const LaunchCfg &cfg) {
    const uint64_t nb = input_mark_blocks(n);
    if (!nb) return;
    if (plain) count_marks_kernel<true><<<(unsigned) nb, kMarkThreads, 0, s>>>(text, n, block_cnt);
    else count_marks_kernel<false><<<(unsigned) nb, kMarkThreads, 0, s>>>(text, n, block_cnt);
    if (cfg.launches) *cfg.launches += 1;
}

void launch_write_marks(const uint8_t *text, uint64_t n, bool plain, const uint64_t *block_off, uint32_t lines_per_record,
                        uint64_t *rec_start, uint64_t *rec_end, uint64_t n_cand, cudaStream_t s, const LaunchCfg &cfg) {
    const uint64_t nb = input_mark_blocks(n);
    if (!nb) return;
    if (plain) write_marks_kernel<true><<<(unsigned) nb, kMarkThreads, 0, s>>>(text, n, block_off, 1, rec_start, rec_end, n_cand);
    else write_marks_kernel<false><<<(unsigned) nb, kMarkThreads, 0, s>>>(text, n, block_off, lines_per_record, rec_start, rec_end, n_cand);
    if (cfg.launches) *cfg.launches += 1;
}

// sc must hold {0xFFFFFFFF, 0xFFFFFFFF, 0, ...} before the call; n_marks = number of marks in the file
void launch_scan_records(const uint8_t *text, uint64_t n, bool plain, const uint64_t *rec_start, const uint64_t *rec_end,
                         uint64_t n_marks, uint32_t lines_per_record, uint32_t n_cand, int trim_left, int trim_right, int rna,
                         int str_threshold, void *info, void *scalars, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_cand) return;
    const int grid = grid_for((uint64_t) n_cand * 32, 256, cfg);
    if (plain)
        scan_records_kernel<true><<<grid, 256, 0, s>>>(text, n, rec_start, rec_end, n_marks, 1, n_cand, trim_left, trim_right, rna,
                                                        str_threshold, (RecInfo *) info, (InputScalars *) scalars);
    else
        scan_records_kernel<false><<<grid, 256, 0, s>>>(text, n, rec_start, rec_end, n_marks, lines_per_record, n_cand, trim_left,
                                                         trim_right, rna, str_threshold, (RecInfo *) info, (InputScalars *) scalars);
    if (cfg.launches) *cfg.launches += 1;
}

void launch_record_totals(const void *info, uint32_t n_rec, void *scalars, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_rec) return;
    record_totals_kernel<<<grid_for(n_rec, 256, cfg), 256, 0, s>>>((const RecInfo *) info, n_rec, (InputScalars *) scalars);
    if (cfg.launches) *cfg.launches += 1;
}

void launch_pack_records(const uint8_t *text, const void *info, uint32_t n_rec, int rna, uint32_t id_step, uint32_t id_first,
                         uint32_t stride, uint32_t *words, uint32_t *len_out, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_rec) return;
    pack_records_kernel<<<grid_for((uint64_t) n_rec * 32, 256, cfg), 256, 0, s>>>(text, (const RecInfo *) info, n_rec, rna, id_step,
                                                                                 id_first, stride, words, len_out);
    if (cfg.launches) *cfg.launches += 1;
}

// A few bytes device -> page-locked host memory without the DMA engine (control read-backs while that engine is busy)
__global__ void peek_kernel(const uint8_t *__restrict__ src, volatile uint8_t *__restrict__ dst, uint32_t bytes) {
    for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
void launch_peek(void *dst_host, const void *src_dev, uint32_t bytes, cudaStream_t s, const LaunchCfg &cfg) {
    if (!bytes) return;
    peek_kernel<<<1, 64, 0, s>>>((const uint8_t *) src_dev, (volatile uint8_t *) dst_host, bytes);
    if (cfg.launches) *cfg.launches += 1;
}

// src: device memory, dst: page-locked host memory (both 16-byte aligned)
void launch_copy_to_host(void *dst_host, const void *src_dev, uint64_t bytes, cudaStream_t s, const LaunchCfg &cfg) {
    if (!bytes) return;
    static const uint64_t max_blocks = getenv("ALGA_FE_COPY_BLOCKS") ? (uint64_t) atoi(getenv("ALGA_FE_COPY_BLOCKS")) : 64;  // tuning knob
    uint64_t blocks = (bytes / 16 + 255) / 256;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    copy_to_host_kernel<<<(unsigned) blocks, 256, 0, s>>>((const uint8_t *) src_dev, (uint8_t *) dst_host, bytes);
    if (cfg.launches) *cfg.launches += 1;
}

// scalars: {max_len, err} zeroed before the call
void launch_remap_flags(const ReadsDev &R, const uint8_t *mask, uint32_t n_units, uint32_t *flag, void *scalars, cudaStream_t s,
                        const LaunchCfg &cfg) {
    if (!n_units) return;
    remap_flags_kernel<<<grid_for(n_units, 256, cfg), 256, 0, s>>>(R, mask, n_units, flag, (RemapScalars *) scalars);
    if (cfg.launches) *cfg.launches += 1;
}

void launch_remap_scatter(const ReadsDev &R, uint32_t n_units, const uint32_t *flag, const uint32_t *pos, uint32_t stride,
                          uint32_t *words, uint32_t *len_out, uint32_t *old_id, uint8_t *paired_offset, uint32_t min_keep_len,
                          cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_units) return;
    remap_scatter_kernel<<<grid_for((uint64_t) n_units * 2 * stride, 256, cfg, 16), 256, 0, s>>>(R, n_units, flag, pos, stride, words,
                                                                                                len_out, old_id, paired_offset, min_keep_len);
    if (cfg.launches) *cfg.launches += 1;
}

}  // namespace alga
