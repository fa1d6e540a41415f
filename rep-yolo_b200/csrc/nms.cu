// non_max_suppression on the GPU, bit-exact against the reference (utils/general.py:953-1045 + torchvision.ops.nms).
//
//   1. nms_compact  : one CTA per image: obj > conf (fp32), conf = obj*cls (nc > 1) or obj (nc == 1, general.py:994-998),
//                     xywh->xyxy (general.py:265-272, same op order), best-class or multi-label rows, class filter;
//                     ordered (prefix-sum) compaction -> rows[n][6] + sort keys.  <true>: driven by the `obj > conf`
//                     ballot words the Detect epilogue left (ry_decode_filter): only the rows that passed are read.
//   2. radix sort   : stable LSD radix sort (4 x 8 bit) of the keys per image, score descending; stability gives the
//                     "ties -> lower index first" order of torchvision's stable sort.  hist -> scan -> scatter per pass.
//   3. nms_scan     : one CTA per image consumes the sorted candidates in steps of 128: suppress by the boxes kept so
//                     far, build the 128x128 bit mask of the survivors (class-offset boxes, general.py:1027-1028), resolve
//                     it with a deterministic serial keep-scan, stop at max_det keeps (exact: general.py:1030-1031
//                     truncates AFTER nms and greedy decisions only depend on higher-ranked boxes) or after max_nms
//                     candidates (general.py:1023-1024, stable variant).
// IoU arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) and the threshold compare is done in
// double, exactly like torchvision's CPU kernel.
#include "nms.cuh"

#include <math.h>

#include "common.cuh"

namespace ry {

namespace {

constexpr int kFilterThreads = 1024;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                           // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;     // 4096 keys per CTA
constexpr float kMaxWh = 4096.0f;                        // general.py:965

__device__ __forceinline__ uint32_t desc_key(float s) {
    uint32_t u = __float_as_uint(s);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;          // ascending order-preserving map of all floats
    return ~u;                                           // descending
}

// ---- block-wide exclusive scan of one int per thread (blockDim.x <= 1024), returns total in *total ----
__device__ __forceinline__ int block_excl_scan(int v, int *warp_sums, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int s = lane < nw ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        warp_sums[lane] = s;                             // inclusive over warps
    }
    __syncthreads();
    const int base = warp > 0 ? warp_sums[warp - 1] : 0;
    *total = warp_sums[nw - 1];
    __syncthreads();
    return base + x - v;
}

// ---- candidate filter (general.py:961-1013), shared by the two front ends below ----
struct Cand {
    float x1, y1, x2, y2, obj, best;
    int bj;
};

// number of output rows of candidate p (0 = filtered out); fills the box / best class on the way
__device__ __forceinline__ int cand_eval(const float *__restrict__ p, int nc, float conf, int multi_label,
                                         const int *__restrict__ classes, int n_classes, Cand &c) {
    c.obj = p[4];
    if (!(c.obj > conf)) return 0;                                             // general.py:962, 978
    const float cx = p[0], cy = p[1], w = p[2], h = p[3];
    c.x1 = __fsub_rn(cx, __fdiv_rn(w, 2.0f));                                  // general.py:268-271
    c.y1 = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
    c.x2 = __fadd_rn(cx, __fdiv_rn(w, 2.0f));
    c.y2 = __fadd_rn(cy, __fdiv_rn(h, 2.0f));
    int cnt = 0;
    if (multi_label) {                                                         // general.py:1004-1006
        for (int j = 0; j < nc; ++j) {
            const float v = __fmul_rn(p[5 + j], c.obj);
            bool ok = v > conf;
            if (ok && n_classes > 0) {
                ok = false;
                for (int q = 0; q < n_classes; ++q) ok |= (classes[q] == j);
            }
            cnt += ok ? 1 : 0;
        }
    } else {                                                                   // general.py:1008-1009
        c.best = 0.0f;
        c.bj = 0;
        for (int j = 0; j < nc; ++j) {
            const float v = (nc == 1) ? c.obj : __fmul_rn(p[5 + j], c.obj);
            if (j == 0 || v > c.best) { c.best = v; c.bj = j; }
        }
        bool ok = c.best > conf;
        if (ok && n_classes > 0) {                                             // general.py:1012-1013
            ok = false;
            for (int q = 0; q < n_classes; ++q) ok |= (classes[q] == c.bj);
        }
        cnt = ok ? 1 : 0;
    }
    return cnt;
}

// writes the rows of a candidate that cand_eval counted (cnt > 0) at ordered position `off`; returns the next position
__device__ __forceinline__ int cand_emit(const float *__restrict__ p, const Cand &c, int nc, float conf, int multi_label,
                                         const int *__restrict__ classes, int n_classes, float *__restrict__ R,
                                         uint32_t *__restrict__ Kb, uint32_t *__restrict__ Ib, int off) {
    if (multi_label) {
        for (int j = 0; j < nc; ++j) {
            const float v = __fmul_rn(p[5 + j], c.obj);
            bool ok = v > conf;
            if (ok && n_classes > 0) {
                ok = false;
                for (int q = 0; q < n_classes; ++q) ok |= (classes[q] == j);
            }
            if (ok) {
                float *r = R + (size_t)off * 6;
                r[0] = c.x1; r[1] = c.y1; r[2] = c.x2; r[3] = c.y2; r[4] = v; r[5] = (float)j;
                Kb[off] = desc_key(v);
                Ib[off] = (uint32_t)off;
                ++off;
            }
        }
    } else {
        float *r = R + (size_t)off * 6;
        r[0] = c.x1; r[1] = c.y1; r[2] = c.x2; r[3] = c.y2; r[4] = c.best; r[5] = (float)c.bj;
        Kb[off] = desc_key(c.best);
        Ib[off] = (uint32_t)off;
        ++off;
    }
    return off;
}

// Ordered compaction, one CTA per image.  Work unit = one warp x 32 consecutive candidates (a "word"):
//   phase 1  every warp walks its words (stride = warps per CTA, no block barrier inside): lane l evaluates candidate 32w + l,
//            the warp's row count and the ballot of the passing lanes go to shared memory;
//   phase 2  ONE block-wide exclusive scan over the per-word counts (ordered, deterministic: no atomics);
//   phase 3  the warps revisit their words, only the passing lanes re-read their row, and write rows / sort keys at
//            word offset + intra-warp prefix.
// MASKED = false (ry_nms): every row of pred is read and tested.  MASKED = true (ry_nms_filtered): the Detect epilogue
// already evaluated `obj > conf` for every candidate it decoded and left the result as ballot words
// mask[b][ceil(N / 32)] (ry_decode_filter); rows whose bit is clear are never read.
template <bool MASKED>
__global__ void __launch_bounds__(kFilterThreads) nms_compact_kernel(const float *__restrict__ pred, const uint32_t *__restrict__ mask,
                                                                     int N, int nc, float conf, int multi_label,
                                                                     const int *__restrict__ classes, int n_classes, size_t cap,
                                                                     float *__restrict__ rows, uint32_t *__restrict__ keys,
                                                                     uint32_t *__restrict__ idx, int *__restrict__ counts) {
    extern __shared__ int sm_words[];                     // [words] row count -> exclusive offset | [words] ballot of passing lanes
    __shared__ int warp_sums[32];
    const int b = blockIdx.x, no = 5 + nc, words = (N + 31) >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    int *wcnt = sm_words;
    uint32_t *wpass = reinterpret_cast<uint32_t *>(sm_words + words);
    const float *P = pred + (size_t)b * N * no;
    const uint32_t *M = MASKED ? mask + (size_t)b * words : nullptr;
    float *R = rows + (size_t)b * cap * 6;
    uint32_t *Kb = keys + (size_t)b * cap, *Ib = idx + (size_t)b * cap;
    // kU words per trip: the objectness loads of all of them are in flight before the first test (a warp walks ~25 words of a
    // 25200-candidate image; one dependent load per trip made this phase pure memory latency)
    constexpr int kU = 4;
    for (int w0 = warp; w0 < words; w0 += kU * nwarps) {
        float obj[kU];
        uint32_t bits[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int w = w0 + u * nwarps, i = 32 * w + lane;
            bits[u] = w < words ? (MASKED ? M[w] : 0xffffffffu) : 0u;
            obj[u] = (i < N && ((bits[u] >> lane) & 1u)) ? P[(size_t)i * no + 4] : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int w = w0 + u * nwarps, i = 32 * w + lane;
            if (w >= words) break;
            Cand c;
            int cnt = 0;
            if (obj[u] > conf) cnt = cand_eval(P + (size_t)i * no, nc, conf, multi_label, classes, n_classes, c);
            const uint32_t pm = __ballot_sync(0xffffffffu, cnt > 0);
            const int tot = multi_label ? __reduce_add_sync(0xffffffffu, cnt) : __popc(pm);
            if (lane == 0) { wcnt[w] = tot; wpass[w] = pm; }
        }
    }
    __syncthreads();
    int base = 0;
    for (int w0 = 0; w0 < words; w0 += blockDim.x) {
        const int w = w0 + threadIdx.x;
        const int v = w < words ? wcnt[w] : 0;
        int total;
        const int ex = block_excl_scan(v, warp_sums, &total);
        if (w < words) wcnt[w] = base + ex;
        base += total;
    }
    __syncthreads();
    for (int w = warp; w < words; w += nwarps) {
        const uint32_t pm = wpass[w];
        if (!pm) continue;
        const int i = 32 * w + lane;
        const float *p = P + (size_t)i * no;
        Cand c;
        int cnt = 0;
        if ((pm >> lane) & 1u) cnt = cand_eval(p, nc, conf, multi_label, classes, n_classes, c);
        int pre;
        if (multi_label) {                               // rows per candidate vary: shuffle scan
            int x = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            pre = x - cnt;
        } else {
            pre = __popc(pm & ((1u << lane) - 1u));
        }
        if (cnt > 0) cand_emit(p, c, nc, conf, multi_label, classes, n_classes, R, Kb, Ib, wcnt[w] + pre);
    }
    if (threadIdx.x == 0) counts[b] = base;
}

// ---- radix sort ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint32_t *__restrict__ keys, const int *__restrict__ counts,
                                                                 size_t cap, int shift, int nblk, uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[256];
    const int b = blockIdx.y, blk = blockIdx.x;
    const int n = counts[b];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t *K = keys + (size_t)b * cap;
    const int start = blk * kSortTile;
    for (int i = start + threadIdx.x; i < min(start + kSortTile, n); i += kSortThreads) atomicAdd(&h[(K[i] >> shift) & 255], 1u);
    __syncthreads();
    hist[((size_t)b * 256 + threadIdx.x) * nblk + blk] = h[threadIdx.x];
}

// exclusive scan of hist[b][digit][blk] in (digit-major, blk-minor) order -> global output offsets
__global__ void __launch_bounds__(256) sort_scan_kernel(uint32_t *__restrict__ hist, int nblk) {
    __shared__ int warp_sums[32];
    const int b = blockIdx.x, d = threadIdx.x;
    uint32_t *h = hist + ((size_t)b * 256 + d) * nblk;
    int sum = 0;
    for (int j = 0; j < nblk; ++j) sum += (int)h[j];
    int total;
    int run = block_excl_scan(sum, warp_sums, &total);
    for (int j = 0; j < nblk; ++j) {
        const int c = (int)h[j];
        h[j] = (uint32_t)run;
        run += c;
    }
}

__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ idx_in,
                                                                    uint32_t *__restrict__ keys_out, uint32_t *__restrict__ idx_out,
                                                                    const int *__restrict__ counts, size_t cap, int shift, int nblk,
                                                                    const uint32_t *__restrict__ hist) {
    __shared__ uint32_t cnt[kSortWarps][256];
    const int b = blockIdx.y, blk = blockIdx.x;
    const int n = counts[b];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t *K = keys_in + (size_t)b * cap, *I = idx_in + (size_t)b * cap;
    const int wstart = blk * kSortTile + warp * (32 * kSortItems);   // each warp owns a contiguous run of the tile
    uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = wstart + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        rank[r] = 0;
        if (valid) {
            key[r] = K[i];
            val[r] = I[i];
            const uint32_t d = (key[r] >> shift) & 255;
            const uint32_t peers = __match_any_sync(vmask, d);
            const uint32_t prior = cnt[warp][d];
            __syncwarp(vmask);
            if ((peers & ((1u << lane) - 1)) == 0) cnt[warp][d] = prior + __popc(peers);   // lowest lane of the group
            __syncwarp(vmask);
            rank[r] = prior + __popc(peers & ((1u << lane) - 1));
        }
    }
    __syncthreads();
    {   // per digit: exclusive prefix over warps + global base of this (image, digit, block)
        const int d = threadIdx.x;
        uint32_t run = hist[((size_t)b * 256 + d) * nblk + blk];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    uint32_t *Ko = keys_out + (size_t)b * cap, *Io = idx_out + (size_t)b * cap;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = wstart + r * 32 + lane;
        if (i < n) {
            const uint32_t pos = cnt[warp][(key[r] >> shift) & 255] + rank[r];
            Ko[pos] = key[r];
            Io[pos] = val[r];
        }
    }
}

// ---- one-kernel sort for small candidate sets -------------------------------------------------------------------
// When an image has at most a few tens of thousands of candidates (every single-label configuration: cap = N), the twelve
// launches of the multi-CTA sort above cost more in launch / drain latency than the sort itself.  Here ONE CTA per image runs
// all four stable LSD passes over the image's key / index arrays (global memory, a few hundred KB per image: L2 resident):
// histogram (shared atomics) -> digit bases -> tile-by-tile ordered scatter (same warp-ordered multi-split as above).
constexpr int kSoloThreads = 1024;
constexpr int kSoloWarps = kSoloThreads / 32;
constexpr int kSoloItems = 8;                            // keys per thread per tile
constexpr int kSoloTile = kSoloThreads * kSoloItems;
constexpr int kSoloMaxCap = 32768;

__global__ void __launch_bounds__(kSoloThreads) sort_image_kernel(uint32_t *__restrict__ keys0, uint32_t *__restrict__ idx0,
                                                                   uint32_t *__restrict__ keys1, uint32_t *__restrict__ idx1,
                                                                   const int *__restrict__ counts, size_t cap) {
    __shared__ uint32_t cnt[kSoloWarps][256];            // per-warp digit counts of the current tile -> output positions
    __shared__ uint32_t dbase[256];                      // running output position of each digit
    __shared__ int warp_sums[32];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = min(counts[b], (int)cap);
    uint32_t *Ka = keys0 + (size_t)b * cap, *Kb = keys1 + (size_t)b * cap;
    uint32_t *Ia = idx0 + (size_t)b * cap, *Ib = idx1 + (size_t)b * cap;
    __shared__ int skip_pass;
    int flips = 0;                                       // executed passes so far: the live copy is in buffer (flips & 1)
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        const uint32_t *Ks = (flips & 1) ? Kb : Ka, *Is = (flips & 1) ? Ib : Ia;   // executed passes ping-pong: a -> b -> a ...
        uint32_t *Kd = (flips & 1) ? Ka : Kb, *Id = (flips & 1) ? Ia : Ib;
        if (tid < 256) dbase[tid] = 0;
        if (tid == 0) skip_pass = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kSoloThreads) atomicAdd(&dbase[(Ks[i] >> shift) & 255], 1u);
        __syncthreads();
        // every key has the same digit (scores of one binade share the top byte; tied scores share all four): the stable
        // scatter of this pass would be the identity -> skip it
        if (tid < 256 && dbase[tid] == (uint32_t)n && n > 0) skip_pass = 1;
        __syncthreads();
        const bool skip = skip_pass != 0 || n == 0;
        __syncthreads();                                 // everybody has read the flag before the next pass resets it
        if (skip) continue;
        {   // exclusive scan of the 256 digit counts (every thread takes part in the block scan; threads >= 256 add zero)
            const int c = tid < 256 ? (int)dbase[tid] : 0;
            int total;
            const int ex = block_excl_scan(c, warp_sums, &total);
            if (tid < 256) dbase[tid] = (uint32_t)ex;
        }
        __syncthreads();
        for (int t0 = 0; t0 < n; t0 += kSoloTile) {
            for (int i = tid; i < kSoloWarps * 256; i += kSoloThreads) (&cnt[0][0])[i] = 0;
            __syncthreads();
            const int wstart = t0 + warp * (32 * kSoloItems);    // each warp owns a contiguous run of the tile
            uint32_t key[kSoloItems], val[kSoloItems], rank[kSoloItems];
#pragma unroll
            for (int r = 0; r < kSoloItems; ++r) {
                const int i = wstart + r * 32 + lane;
                const bool valid = i < n;
                const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
                rank[r] = 0; key[r] = 0; val[r] = 0;
                if (valid) {
                    key[r] = Ks[i];
                    val[r] = Is[i];
                    const uint32_t d = (key[r] >> shift) & 255;
                    const uint32_t peers = __match_any_sync(vmask, d);
                    const uint32_t prior = cnt[warp][d];
                    __syncwarp(vmask);
                    if ((peers & ((1u << lane) - 1)) == 0) cnt[warp][d] = prior + __popc(peers);   // lowest lane of the group
                    __syncwarp(vmask);
                    rank[r] = prior + __popc(peers & ((1u << lane) - 1));
                }
            }
            __syncthreads();
            if (tid < 256) {                                     // per digit: exclusive prefix over warps + running base
                uint32_t run = dbase[tid];
#pragma unroll 8
                for (int w = 0; w < kSoloWarps; ++w) {
                    const uint32_t c = cnt[w][tid];
                    cnt[w][tid] = run;
                    run += c;
                }
                dbase[tid] = run;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kSoloItems; ++r) {
                const int i = wstart + r * 32 + lane;
                if (i < n) {
                    const uint32_t pos = cnt[warp][(key[r] >> shift) & 255] + rank[r];
                    Kd[pos] = key[r];
                    Id[pos] = val[r];
                }
            }
            __syncthreads();
        }
        ++flips;
    }
    if (flips & 1) {                                     // an odd number of passes ran: the result is in b, the scan reads a
        for (int i = tid; i < n; i += kSoloThreads) { Ka[i] = Kb[i]; Ia[i] = Ib[i]; }
    }
}

// ---- greedy scan ---------------------------------------------------------------------------------------------------
// torchvision's CPU kernel suppresses iff (double)ovr > (double)thr with ovr the fp32 quotient.  For a float q,
// (double)q > thr  <=>  q >= t_up, t_up = the smallest float whose value exceeds thr (computed on the host), so the
// compare runs in fp32; and when the boxes do not intersect (inter == 0) the quotient is 0, -0 or NaN, never > thr >= 0,
// so the IEEE division is only paid for overlapping pairs.  `nonneg` = (thr >= 0).
struct IouThr { float t_up; int nonneg; };

__device__ __forceinline__ bool iou_gt(float ax1, float ay1, float ax2, float ay2, float aa, float bx1, float by1, float bx2,
                                       float by2, float ba, IouThr thr) {
    const float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1), xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
    const float w = fmaxf(__fsub_rn(xx2, xx1), 0.0f), h = fmaxf(__fsub_rn(yy2, yy1), 0.0f);
    const float inter = __fmul_rn(w, h);
    if (inter == 0.0f && thr.nonneg) return false;
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ba), inter));
    return ovr >= thr.t_up;
}

// One CTA per image, steps of kCh = 128 sorted candidates (round 1: the 512-candidate step spent its time in a 512 x 512
// pair mask and a ~100-cycle-per-keep warp scan on ONE SM per image, although max_det = 300 keeps are usually found within
// the first few hundred candidates).  Per step:
//   1. kept-list test: kSub = 4 threads per candidate share the (<= max_det) kept boxes, 4 independent IoUs per trip each;
//   2. ballot-compaction of the survivors (order preserved);
//   3. survivor x survivor mask, one 32-bit word per thread (128 x 4 words);
//   4. keep-scan by ONE thread with the 128 dead bits in four registers: per keep one 16-byte shared-memory row read;
//   5. publish the new keeps (kept-box list + output rows).
// The rows of the NEXT step are fetched (sort order -> row gather, two dependent loads) while the current step runs.
constexpr int kScanThreads = 512;
constexpr int kCh = 128;
constexpr int kSub = kScanThreads / kCh;

__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(const float *__restrict__ rows, const uint32_t *__restrict__ order,
                                                                const int *__restrict__ counts, size_t cap, IouThr iou_thr,
                                                                int agnostic, int max_det, int max_nms, float *__restrict__ out,
                                                                int *__restrict__ out_counts) {
    extern __shared__ __align__(16) unsigned char smraw[];
    // survivor mask rows (16 B each) | kept boxes | survivor boxes of the step (SoA) | kept areas | control
    uint4 *mask = reinterpret_cast<uint4 *>(smraw);                                          // [kCh] bit j of row i: i suppresses j (j > i)
    const int kcap = (max_det + 3) & ~3;                                                     // kept list padded to the unroll of the test loop
    float4 *kbox = reinterpret_cast<float4 *>(mask + kCh);                                   // [kcap] (x1, y1, x2, y2) with class offset
    float *cx1 = reinterpret_cast<float *>(kbox + kcap), *cy1 = cx1 + kCh, *cx2 = cy1 + kCh, *cy2 = cx2 + kCh, *car = cy2 + kCh;
    float *kar = car + kCh;                                                                  // [kcap] areas
    int *ctl = reinterpret_cast<int *>(kar + kcap);                                          // [0] kept so far, [1] new keeps of the step
    int *newkeep = ctl + 2;                                                                  // [kCh] survivor indices kept
    float *srow = reinterpret_cast<float *>(newkeep + kCh);                                  // [kCh][6] survivor rows as they go out
    int *wcnt = reinterpret_cast<int *>(srow + kCh * 6);                                     // [kScanThreads / 32] survivors per warp

    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cand = tid / kSub, sub = tid - cand * kSub;
    const int n = min(counts[b], max_nms);
    const float *R = rows + (size_t)b * cap * 6;
    const uint32_t *O = order + (size_t)b * cap;
    float *outb = out + (size_t)b * max_det * 6;
    if (tid == 0) { ctl[0] = 0; ctl[1] = 0; }
    for (int k = tid; k < kcap; k += kScanThreads) {             // padding entries never intersect anything (inter == 0 -> not suppressed)
        kbox[k] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f);
        kar[k] = 0.0f;
    }
    // rows of the first step (the leader lane of each candidate loads, its kSub - 1 siblings get the box by shuffle)
    float nr0 = 0, nr1 = 0, nr2 = 0, nr3 = 0, nr4 = 0, nr5 = 0;
    if (sub == 0 && cand < n) {
        const float *r = R + (size_t)O[cand] * 6;
        nr0 = r[0]; nr1 = r[1]; nr2 = r[2]; nr3 = r[3]; nr4 = r[4]; nr5 = r[5];
    }
    __syncthreads();

    for (int c0 = 0; c0 < n; c0 += kCh) {
        const int cs = min(kCh, n - c0);
        const int kept0 = ctl[0];
        if (kept0 >= max_det) break;
        const float r0 = nr0, r1 = nr1, r2 = nr2, r3 = nr3, r4 = nr4, r5 = nr5;
        if (sub == 0 && c0 + kCh + cand < n) {                   // next step's rows: in flight during this step
            const float *r = R + (size_t)O[c0 + kCh + cand] * 6;
            nr0 = r[0]; nr1 = r[1]; nr2 = r[2]; nr3 = r[3]; nr4 = r[4]; nr5 = r[5];
        }
        // ---- my box (class offset added in fp32 BEFORE the IoU, general.py:1027-1028); suppressed by an earlier keep? ----
        float x1 = 0, y1 = 0, x2 = 0, y2 = 0, ar = 0;
        if (sub == 0) {
            const float off = agnostic ? __fmul_rn(r5, 0.0f) : __fmul_rn(r5, kMaxWh);
            x1 = __fadd_rn(r0, off); y1 = __fadd_rn(r1, off); x2 = __fadd_rn(r2, off); y2 = __fadd_rn(r3, off);
            ar = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
        }
        const int src = lane & ~(kSub - 1);
        x1 = __shfl_sync(0xffffffffu, x1, src); y1 = __shfl_sync(0xffffffffu, y1, src);
        x2 = __shfl_sync(0xffffffffu, x2, src); y2 = __shfl_sync(0xffffffffu, y2, src);
        ar = __shfl_sync(0xffffffffu, ar, src);
        bool sup = false;
        if (cand < cs) {
            for (int k = 4 * sub; k < kept0 && !sup; k += 4 * kSub) {          // four independent tests per trip (the loop is latency bound)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 kb = kbox[k + u];
                    sup |= (k + u < kept0) & iou_gt(kb.x, kb.y, kb.z, kb.w, kar[k + u], x1, y1, x2, y2, ar, iou_thr);
                }
            }
        }
        {
            int s = sup ? 1 : 0;
#pragma unroll
            for (int o = 1; o < kSub; o <<= 1) s |= __shfl_xor_sync(0xffffffffu, s, o);
            sup = s != 0;
        }
        const bool alive = sub == 0 && cand < cs && !sup;
        // ---- compact the survivors (order preserved): only they can be kept or suppress anything from here on ----
        const uint32_t am = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) wcnt[warp] = __popc(am);
        __syncthreads();
        int base = 0, S = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) {
            const int c = wcnt[w];
            if (w < warp) base += c;
            S += c;
        }
        if (alive) {
            const int ci = base + __popc(am & ((1u << lane) - 1));
            cx1[ci] = x1; cy1[ci] = y1; cx2[ci] = x2; cy2[ci] = y2; car[ci] = ar;
            float *o = srow + ci * 6;                            // the row as it goes out (un-offset box, general.py:1040)
            o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4; o[5] = r5;
        }
        __syncthreads();
        // ---- survivor mask: thread (i = tid / 4, w = tid % 4) -> bits j = 32w .. 32w+31 (only j > i matters) of row i ----
        {
            const int i = tid >> 2, w = tid & 3;
            uint32_t bits = 0;
            if (i < S && 32 * w + 31 > i) {
                const float ax1 = cx1[i], ay1 = cy1[i], ax2 = cx2[i], ay2 = cy2[i], aa = car[i];
                const int j0 = max(32 * w, i + 1), j1 = min(32 * w + 32, S);
                for (int j = j0; j < j1; ++j)
                    if (iou_gt(ax1, ay1, ax2, ay2, aa, cx1[j], cy1[j], cx2[j], cy2[j], car[j], iou_thr)) bits |= 1u << (j & 31);
            }
            reinterpret_cast<uint32_t *>(mask)[tid] = bits;
        }
        __syncthreads();
        // ---- deterministic keep-scan, one thread, dead set in registers ----
        if (tid == 0) {
            uint32_t d0, d1, d2, d3;
            {
                auto tail = [&](int wbase) -> uint32_t { return S >= wbase + 32 ? 0u : (S <= wbase ? 0xffffffffu : ~((1u << (S - wbase)) - 1u)); };
                d0 = tail(0); d1 = tail(32); d2 = tail(64); d3 = tail(96);
            }
            int kept = kept0, nk = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                while (kept < max_det) {
                    const uint32_t dw = w == 0 ? d0 : (w == 1 ? d1 : (w == 2 ? d2 : d3));
                    const uint32_t al = ~dw;
                    if (!al) break;
                    const int bit = __ffs(al) - 1, i = 32 * w + bit;
                    newkeep[nk++] = i;
                    ++kept;
                    const uint4 m = mask[i];
                    d0 |= m.x; d1 |= m.y; d2 |= m.z; d3 |= m.w;
                    if (w == 0) d0 |= 1u << bit; else if (w == 1) d1 |= 1u << bit; else if (w == 2) d2 |= 1u << bit; else d3 |= 1u << bit;   // consumed
                }
            }
            ctl[0] = kept;
            ctl[1] = nk;
        }
        __syncthreads();
        // ---- publish the new keeps: kept-box list (for later steps) and output rows (un-offset boxes, general.py:1040) ----
        const int nk = ctl[1];
        for (int t = tid; t < nk; t += kScanThreads) {
            const int i = newkeep[t], slot = kept0 + t;
            kbox[slot] = make_float4(cx1[i], cy1[i], cx2[i], cy2[i]); kar[slot] = car[i];
            const float *r = srow + i * 6;
            float *o = outb + (size_t)slot * 6;
#pragma unroll
            for (int q = 0; q < 6; ++q) o[q] = r[q];
        }
        __syncthreads();
    }
    if (tid == 0) out_counts[b] = ctl[0];
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct NmsLayout {
    size_t cap, rows, keys0, keys1, idx0, idx1, counts, hist, classes, total;
    int nblk;
};

NmsLayout nms_layout(int B, int N, int nc, int multi_label) {
    NmsLayout L;
    L.cap = (size_t)N * ((multi_label && nc > 1) ? nc : 1);
    L.nblk = (int)((L.cap + kSortTile - 1) / kSortTile);
    size_t o = 0;
    L.rows = o;    o = align_up(o + (size_t)B * L.cap * 6 * 4, 256);
    L.keys0 = o;   o = align_up(o + (size_t)B * L.cap * 4, 256);
    L.keys1 = o;   o = align_up(o + (size_t)B * L.cap * 4, 256);
    L.idx0 = o;    o = align_up(o + (size_t)B * L.cap * 4, 256);
    L.idx1 = o;    o = align_up(o + (size_t)B * L.cap * 4, 256);
    L.counts = o;  o = align_up(o + (size_t)B * 4, 256);
    L.hist = o;    o = align_up(o + (size_t)B * 256 * L.nblk * 4, 256);
    L.classes = o; o = align_up(o + 1024 * 4, 256);
    L.total = o;
    return L;
}

}  // namespace

size_t nms_workspace_bytes(int B, int N, int nc, int multi_label) { return nms_layout(B, N, nc, multi_label).total; }

int nms_launch_count(int B, int N, int nc, int multi_label) {
    (void)B;
    const size_t cap = (size_t)N * ((multi_label && nc > 1) ? nc : 1);
    return cap <= (size_t)kSoloMaxCap ? 3 : 1 + 4 * 3 + 1;        // filter, sort (one kernel, or hist / scan / scatter x 4), scan
}

int nms_run(const float *pred, const uint32_t *cand_mask, int B, int N, int nc, float conf, double iou, const int32_t *classes_host,
            int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float *out, int32_t *counts, void *workspace,
            size_t workspace_bytes, cudaStream_t st) {
    multi_label = (multi_label && nc > 1) ? 1 : 0;                                         // general.py:970
    const NmsLayout L = nms_layout(B, N, nc, multi_label);
    if (workspace_bytes < L.total) RY_FAIL("ry_nms: workspace too small");
    if (n_classes > 1024) RY_FAIL("ry_nms: more than 1024 classes in the filter list");
    if (max_det < 1 || max_det > 4096) RY_FAIL("ry_nms: max_det out of range");
    if (B <= 0 || N <= 0) RY_FAIL("ry_nms: empty input");
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    float *rows = reinterpret_cast<float *>(ws + L.rows);
    uint32_t *k0 = reinterpret_cast<uint32_t *>(ws + L.keys0), *k1 = reinterpret_cast<uint32_t *>(ws + L.keys1);
    uint32_t *i0 = reinterpret_cast<uint32_t *>(ws + L.idx0), *i1 = reinterpret_cast<uint32_t *>(ws + L.idx1);
    int *cnt = reinterpret_cast<int *>(ws + L.counts);
    uint32_t *hist = reinterpret_cast<uint32_t *>(ws + L.hist);
    int *cls = reinterpret_cast<int *>(ws + L.classes);
    if (n_classes > 0) RY_CUDA(cudaMemcpyAsync(cls, classes_host, (size_t)n_classes * 4, cudaMemcpyHostToDevice, st));

    const size_t csm = (size_t)((N + 31) / 32) * 8;
    if (csm > 200 * 1024) RY_FAIL("ry_nms: more than 800k candidates per image");
    if (cand_mask != nullptr) {
        RY_CUDA(RY_ENSURE_DYN_SMEM(nms_compact_kernel<true>, 200 * 1024));
        nms_compact_kernel<true><<<B, kFilterThreads, csm, st>>>(pred, cand_mask, N, nc, conf, multi_label, cls, n_classes, L.cap, rows, k0, i0, cnt);
    } else {
        RY_CUDA(RY_ENSURE_DYN_SMEM(nms_compact_kernel<false>, 200 * 1024));
        nms_compact_kernel<false><<<B, kFilterThreads, csm, st>>>(pred, nullptr, N, nc, conf, multi_label, cls, n_classes, L.cap, rows, k0, i0, cnt);
    }
    static const bool no_solo = getenv("RY_NMS_MULTI_SORT") != nullptr;
    const bool solo = !no_solo && L.cap <= (size_t)kSoloMaxCap;
    if (solo) sort_image_kernel<<<B, kSoloThreads, 0, st>>>(k0, i0, k1, i1, cnt, L.cap);      // four passes: the result is back in k0 / i0
    const dim3 sgrid(L.nblk, B);
    for (int pass = 0; pass < 4 && !solo; ++pass) {
        const int shift = 8 * pass;
        sort_hist_kernel<<<sgrid, kSortThreads, 0, st>>>(k0, cnt, L.cap, shift, L.nblk, hist);
        sort_scan_kernel<<<B, 256, 0, st>>>(hist, L.nblk);
        sort_scatter_kernel<<<sgrid, kSortThreads, 0, st>>>(k0, i0, k1, i1, cnt, L.cap, shift, L.nblk, hist);
        uint32_t *t = k0; k0 = k1; k1 = t;
        t = i0; i0 = i1; i1 = t;
    }
    const size_t smem = (size_t)kCh * 16 + (size_t)5 * ((max_det + 3) & ~3) * 4 + (size_t)5 * kCh * 4 + 2 * 4 + (size_t)7 * kCh * 4 +
                        (kScanThreads / 32) * 4;
    RY_CUDA(RY_ENSURE_DYN_SMEM(nms_scan_kernel, 200 * 1024));
    IouThr thr;
    {   // smallest float strictly greater than the double threshold
        float tf = (float)iou;
        if (!((double)tf > iou)) tf = nextafterf(tf, INFINITY);
        else while ((double)nextafterf(tf, -INFINITY) > iou) tf = nextafterf(tf, -INFINITY);
        thr.t_up = tf;
        thr.nonneg = iou >= 0.0 ? 1 : 0;
    }
    nms_scan_kernel<<<B, kScanThreads, smem, st>>>(rows, i0, cnt, L.cap, thr, agnostic, max_det, max_nms, out, counts);
    RY_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ry
