// Plan / executor and the C ABI (include/repyolo_b200.h).
//
// A plan is the fused deploy graph of the reference's Model.fuse() + forward_once (models/yolo.py:569-619, 681-704) as a
// flat op list over NHWC bf16 tensors.  ry_plan_create packs the weights once (bf16, K-major, padded to the UMMA K block);
// ry_plan_bind lays the tensors out in the caller's workspace and encodes the TMA descriptors for one input shape;
// ry_forward / ry_run_ops only launch kernels on the caller's stream.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/repyolo_b200.h"
#include "common.cuh"
#include "conv_umma.cuh"
#include "conv_chain.cuh"
#include "memops.cuh"
#include "nms.cuh"

namespace ry {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }

int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev & 63;
    int n = cache[slot].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[slot].store(n, std::memory_order_relaxed);
    }
    return n;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Tensor {
    ry_tensor_desc d;
    size_t offset = 0, bytes = 0;
    int h = 0, w = 0;
};

struct ConvPacked {          // shape-independent part of a CONV / DETECT op
    int kb = 0, cblk = 0, ksteps_last = 0, ntaps = 0, BN = 0, n_ntiles = 0, k_pad = 0, cout_pad = 0;
    size_t w_dev = 0, b_dev = 0;   // byte offsets in the device weight buffer
    int n_src = 1, src_len[3] = {0, 0, 0};   // multi-source 1x1: each source padded to whole K blocks
};

struct Op {
    ry_op_desc d;
    ConvPacked cp;
    size_t dev[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // device weight-buffer offsets of the op's parameter arrays
    // shape-dependent (filled by bind)
    ConvArgs ca;
    ChainArgs ch;
    size_t post_w_dev[2] = {0, 0}, post_b_dev[2] = {0, 0};   // CONV_CHAIN: pre-swizzled 1x1 weight images / padded biases
    int post_kb[2] = {0, 0}, post_n[2] = {0, 0}, post_wbytes[2] = {0, 0};
    int grid = 0;
    int tmap_first = -1, n_amaps = 1;
    int launches = 0;
};

}  // namespace
}  // namespace ry

struct ry_plan {
    int device = 0, nc = 1;
    std::vector<ry::Tensor> tensors;
    std::vector<ry::Op> ops;
    unsigned char *d_weights = nullptr;
    size_t weight_bytes = 0;
    // bound state
    int B = 0, H = 0, W = 0;
    unsigned char *ws = nullptr;
    size_t ws_bytes = 0, scratch_off = 0;
    CUtensorMap *d_tmaps = nullptr;
    size_t tmaps_cap = 0;
    int n_cand = 0;
    // optional per-op CUDA-event timing (bench.py roofline): events[2*i], events[2*i+1] bracket op i
    int image_u8 = 0;            // ry_plan_set_image_dtype: the `image` pointer is uint8 NCHW (0..255), /255 fused in the stem
    bool profiling = false;
    uint32_t *filter_mask = nullptr;   // set for the duration of ry_decode_filter
    float filter_conf = 0.0f;
    int energies_ready_op = -1;      // index of the VERTICAL op whose energies the preceding fused criss-cross column pass wrote
    std::vector<cudaEvent_t> events;
};

namespace ry {
namespace {

struct Blob {                 // host staging of the device weight buffer
    std::vector<unsigned char> data;
    size_t add(const void *src, size_t bytes) {
        const size_t off = align_up(data.size(), 256);
        data.resize(off + bytes);
        if (src) memcpy(data.data() + off, src, bytes);
        return off;
    }
};

int pack_conv(const ry_op_desc &d, const unsigned char *host, size_t host_bytes, Blob &blob, ConvPacked &cp) {
    const int cin = d.cin, cout = d.cout, k = d.ksize;
    if (k != 1 && k != 3) RY_FAIL("conv: ksize must be 1 or 3");
    if (cin % 8 != 0) RY_FAIL("conv: cin must be a multiple of 8");
    if (d.w_off < 0 || d.b_off < 0 || (size_t)d.w_off + (size_t)cout * cin * k * k * 4 > host_bytes ||
        (size_t)d.b_off + (size_t)cout * 4 > host_bytes)
        RY_FAIL("conv: weight offsets out of range");
    // K block = swizzle span; a partial last block is completed by TMA out-of-bounds zero fill (activations) and by the
    // zero padding of the packed weights
    cp.kb = cin > 32 ? 64 : (cin > 16 ? 32 : 16);
    cp.cblk = (cin + cp.kb - 1) / cp.kb;
    cp.ksteps_last = (cin - (cp.cblk - 1) * cp.kb + 15) / 16;
    cp.n_src = d.n_src > 1 ? d.n_src : 1;
    if (cp.n_src > 1) {
        if (k != 1 || cp.n_src > 3) RY_FAIL("conv: concatenated inputs need a 1x1 conv with at most 3 sources");
        const ry_view *vs[3] = {&d.in0, &d.in1, &d.in2};
        int tot = 0, maxlen = 0;
        cp.cblk = 0;
        for (int i = 0; i < cp.n_src; ++i) { cp.src_len[i] = vs[i]->c_len; tot += vs[i]->c_len; maxlen = std::max(maxlen, vs[i]->c_len); }
        if (tot != cin) RY_FAIL("conv: concatenated input views do not add up to cin");
        cp.kb = maxlen > 32 ? 64 : (maxlen > 16 ? 32 : 16);
        for (int i = 0; i < cp.n_src; ++i) cp.cblk += (cp.src_len[i] + cp.kb - 1) / cp.kb;
        if (cp.cblk > kConvMaxSrcBlocks) RY_FAIL("conv: too many K blocks for a multi-source conv");
    }
    cp.ntaps = k * k;
    cp.k_pad = cp.ntaps * cp.cblk * cp.kb;
    const int c16 = (cout + 15) / 16 * 16;
    cp.BN = c16 <= 256 ? c16 : 256;
    if (c16 > 256 && c16 % 256 != 0) cp.BN = 128;
    {   // HBM-bound 256-wide 1x1 layers: two 128-wide N tiles so that four epilogue teams (4 x 128 TMEM columns) work in parallel
        static const double ai_max = getenv("RY_CONV_SPLIT256_AI") ? atof(getenv("RY_CONV_SPLIT256_AI")) : 160.0;
        const double ai = (double)cin * cout * k * k / (double)(cin + cout);      // FLOP per byte of activation traffic
        if (c16 == 256 && k == 1 && d.kind == RY_OP_CONV && d.out1.tensor < 0 && ai < ai_max) cp.BN = 128;
    }
    cp.n_ntiles = (c16 + cp.BN - 1) / cp.BN;
    cp.cout_pad = cp.n_ntiles * cp.BN;
    const float *w = reinterpret_cast<const float *>(host + d.w_off);
    const float *b = reinterpret_cast<const float *>(host + d.b_off);
    std::vector<__nv_bfloat16> wp((size_t)cp.cout_pad * cp.k_pad, __float2bfloat16(0.0f));
    std::vector<int> kcol(cin);                       // packed K column of input channel ci (within a tap)
    if (cp.n_src > 1) {
        int ci = 0, blk = 0;
        for (int s = 0; s < cp.n_src; ++s) {
            for (int c = 0; c < cp.src_len[s]; ++c) kcol[ci++] = blk * cp.kb + c;
            blk += (cp.src_len[s] + cp.kb - 1) / cp.kb;
        }
    } else {
        for (int ci = 0; ci < cin; ++ci) kcol[ci] = ci;
    }
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int t = 0; t < cp.ntaps; ++t)
                wp[(size_t)co * cp.k_pad + (size_t)t * cp.cblk * cp.kb + kcol[ci]] =
                    __float2bfloat16_rn(w[((size_t)co * cin + ci) * cp.ntaps + t]);
    std::vector<float> bp(cp.cout_pad, 0.0f);
    for (int co = 0; co < cout; ++co) bp[co] = b[co];
    cp.w_dev = blob.add(wp.data(), wp.size() * sizeof(__nv_bfloat16));
    cp.b_dev = blob.add(bp.data(), bp.size() * sizeof(float));
    return 0;
}

// CONV_CHAIN 1x1 stage: weights [cout][cin] fp32 -> bf16 image [N][kb] in the K-major 64/128-byte-swizzled layout the
// tcgen05 B operand expects in shared memory (a plain bulk copy brings it in); bias padded with zeros to N.
int pack_post(const unsigned char *host, size_t host_bytes, int64_t w_off, int64_t b_off, int cin, int cout, Blob &blob, Op &op, int i) {
    if (cin % 8 || cout % 8 || cin > 64 || cout > 64) RY_FAIL("conv chain: fused 1x1 stages need cin, cout <= 64 and multiples of 8");
    if (w_off < 0 || b_off < 0 || (size_t)w_off + (size_t)cout * cin * 4 > host_bytes || (size_t)b_off + (size_t)cout * 4 > host_bytes)
        RY_FAIL("conv chain: weight offsets out of range");
    const float *w = reinterpret_cast<const float *>(host + w_off);
    const float *b = reinterpret_cast<const float *>(host + b_off);
    const int kb = cin > 32 ? 64 : 32, rb = kb * 2, N = (cout + 15) / 16 * 16;
    const uint32_t m = rb == 128 ? 7u : 3u;
    std::vector<__nv_bfloat16> img((size_t)N * kb, __float2bfloat16(0.0f));
    for (int n = 0; n < cout; ++n)
        for (int k = 0; k < cin; ++k) {
            uint32_t lin = (uint32_t)n * rb + (uint32_t)(k / 8) * 16;
            lin ^= ((lin >> 7) & m) << 4;
            img[(lin >> 1) + (k % 8)] = __float2bfloat16_rn(w[(size_t)n * cin + k]);
        }
    std::vector<float> bp(N, 0.0f);
    for (int n = 0; n < cout; ++n) bp[n] = b[n];
    op.post_w_dev[i] = blob.add(img.data(), img.size() * sizeof(__nv_bfloat16));
    op.post_b_dev[i] = blob.add(bp.data(), bp.size() * sizeof(float));
    op.post_kb[i] = kb; op.post_n[i] = N; op.post_wbytes[i] = N * rb;
    return 0;
}

int copy_f32(const unsigned char *host, size_t host_bytes, int64_t off, size_t n, Blob &blob, size_t *dev) {
    if (off < 0 || (size_t)off + n * 4 > host_bytes) RY_FAIL("weight offset out of range");
    *dev = blob.add(host + off, n * 4);
    return 0;
}

void pick_tile(int B, int Ho, int Wo, int *tw, int *th, int *tn, bool even = false) {   // even: 2x2 pooling windows stay inside a tile
    long best_tiles = -1;
    int bw = 1, bh = 1, bn = 1;
    for (int w = std::min(Wo, even ? 64 : 128); w >= 1; --w) {
        for (int h = std::min(Ho, 128 / w); h >= 1; --h) {
            if (even && ((w | h) & 1)) continue;
            const int n = std::max(1, std::min(B, 128 / (w * h)));
            const long tiles = (long)cdiv(Wo, w) * cdiv(Ho, h) * cdiv(B, n);
            if (best_tiles < 0 || tiles < best_tiles) { best_tiles = tiles; bw = w; bh = h; bn = n; }
        }
    }
    *tw = bw; *th = bh; *tn = bn;
}

int encode_map(CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
               const cuuint32_t *box, int kb, bool dense = true) {   // kb = inner box extent in bf16 elements: 64/32/16 -> 128/64/32-byte swizzle, 8 -> none
    EncodeTiledFn enc = get_encode();
    if (!enc) RY_FAIL("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = kb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : (kb == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           dense ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,   // channel-slice views: no over-fetch
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RY_FAIL("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

int view_ok(const ry_plan *p, const ry_view &v, bool required) {
    if (v.tensor < 0) return required ? 1 : 0;
    if (v.tensor >= (int)p->tensors.size()) return 1;
    const Tensor &t = p->tensors[v.tensor];
    if (t.d.kind == RY_T_EXTERNAL) return 0;
    if (v.c_off < 0 || v.c_len <= 0 || v.c_off + v.c_len > t.d.channels || v.c_off % 8 != 0) return 1;
    return 0;
}

inline __nv_bfloat16 *bf(ry_plan *p, int t) { return reinterpret_cast<__nv_bfloat16 *>(p->ws + p->tensors[t].offset); }
inline float *f32(ry_plan *p, int t) { return reinterpret_cast<float *>(p->ws + p->tensors[t].offset); }
inline const float *wf(ry_plan *p, size_t off) { return reinterpret_cast<const float *>(p->d_weights + off); }

// Output channels of one N tile -> store segments (power-of-two widths, each with its own swizzled staging layout and TMA
// store map), spread over the two epilogue column groups.
struct OutPiece { int col0, len, chan, slot; };   // slot: 0 = out0's tensor, 1 = out1's tensor (split store into two tensors)

// widths[] holds one key per store map: slot * 1000 + box width
void build_segments(ConvArgs &a, const OutPiece *pieces, int n_pieces, int max_cols, int widths[8], int *n_widths) {
    int total = 0;
    for (int i = 0; i < n_pieces; ++i) total += pieces[i].len;
    struct S { int col0, ncol, chan, slot; };
    std::vector<S> segs;
    // small N: one segment per output piece (dense rows unless the width is a power of two), the two epilogue groups work
    // as teams on alternate tiles; otherwise power-of-two segments spread over two column groups
    static const int teams_max = getenv("RY_CONV_TEAMS_MAX") ? atoi(getenv("RY_CONV_TEAMS_MAX")) : 128;
    a.ep_teams = (total <= std::max(64, teams_max) && 4 * a.BN <= 512 && n_pieces <= kConvMaxSegs) ? 1 : 0;
    if (total <= 64 && n_pieces > kConvMaxSegs) a.ep_teams = 0;
    if (a.ep_teams && total <= 64) {
        for (int i = 0; i < n_pieces; ++i) segs.push_back({pieces[i].col0, pieces[i].len, pieces[i].chan, pieces[i].slot});
    } else {
        int wmax = 8;
        while (wmax * 2 <= max_cols && wmax * 2 <= std::max(8, total / 2)) wmax *= 2;
        for (int i = 0; i < n_pieces; ++i)
            for (int c = 0; c < pieces[i].len;) {
                int w = wmax;
                while (w > pieces[i].len - c) w /= 2;
                segs.push_back({pieces[i].col0 + c, w, pieces[i].chan + c, pieces[i].slot});
                c += w;
            }
        std::stable_sort(segs.begin(), segs.end(), [](const S &x, const S &y) { return x.ncol > y.ncol; });
    }
    *n_widths = 0;
    int load[2] = {0, 0};
    a.nseg[0] = a.nseg[1] = 0;
    for (const S &sg : segs) {
        int wi = 0;
        const int key = sg.slot * 1000 + sg.ncol;
        while (wi < *n_widths && widths[wi] != key) ++wi;
        if (wi == *n_widths) widths[(*n_widths)++] = key;
        for (int g = 0; g < 2; ++g) {
            if (!a.ep_teams && g != (load[1] < load[0] ? 1 : 0)) continue;
            ConvSeg &d = a.seg[g][a.nseg[g]++];
            d.col0 = (int16_t)sg.col0; d.ncol = (int16_t)sg.ncol; d.chan = sg.chan; d.map = (int16_t)wi;
            d.swz = (int16_t)(sg.ncol == 64 ? 7 : (sg.ncol == 32 ? 3 : (sg.ncol == 16 ? 1 : 0)));
            load[g] += sg.ncol;
            if (!a.ep_teams) break;
        }
    }
}

// Fill the shape-dependent ConvArgs + tensor maps of one CONV / DETECT op.
int bind_conv(ry_plan *p, Op &op, std::vector<CUtensorMap> &maps) {
    const ry_op_desc &d = op.d;
    const ConvPacked &cp = op.cp;
    const Tensor &tin = p->tensors[d.in0.tensor];
    const int s = d.stride, k = d.ksize;
    const int Hi = tin.h, Wi = tin.w, Ho = Hi / s, Wo = Wi / s, B = p->B;
    if (s != 1 && !(s == 2 && k == 3)) RY_FAIL("conv: only 1x1 s1, 3x3 s1 and 3x3 s2 are built");
    ConvArgs &a = op.ca;
    memset(&a, 0, sizeof(a));
    a.kb = cp.kb; a.cblk = cp.cblk; a.ntaps = cp.ntaps; a.kblocks = cp.ntaps * cp.cblk; a.ksteps_last = cp.ksteps_last;
    a.BN = cp.BN; a.n_ntiles = cp.n_ntiles; a.cout_pad = cp.cout_pad;
    a.img_w = Wo; a.img_hw = Ho * Wo;
    a.mode = d.kind == RY_OP_DETECT ? 1 : 0;
    a.pool = (d.kind == RY_OP_CONV && k == 1 && d.level_idx == 1) ? 1 : 0;
    if (a.pool && ((Hi | Wi) & 1)) RY_FAIL("conv: fused max-pool needs an even map");
    op.tmap_first = (int)maps.size();
    const size_t esz = 2;
    const cuuint64_t ctot = (cuuint64_t)tin.d.channels;
    __nv_bfloat16 *in_base = bf(p, d.in0.tensor) + d.in0.c_off;
    const bool in_dense = d.in0.c_len == tin.d.channels;
    static const bool no_halo = getenv("RY_CONV_NO_HALO") != nullptr;
    CUtensorMap m;
    a.a_mode = A_BOX;
    static const char *acc_env = getenv("RY_CONV_NACC");
    if (k == 1) {
        const cuuint64_t P = (cuuint64_t)B * Hi * Wi;
        a.tw = 128; a.th = 1; a.tn = 1;
        a.Wo = (int)P; a.Ho = 1; a.Bo = 1;
        if (a.pool) {                                  // 2-D pixel tiles (even extents) so that every 2x2 window lies in one tile
            pick_tile(B, Hi, Wi, &a.tw, &a.th, &a.tn, true);
            a.Wo = Wi; a.Ho = Hi; a.Bo = B;
        }
        const cuuint32_t box[4] = {(cuuint32_t)cp.kb, (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
        a.n_src = cp.n_src;
        const ry_view *vs[3] = {&d.in0, &d.in1, &d.in2};
        int blk = 0;
        for (int si = 0; si < cp.n_src; ++si) {
            const Tensor &ts = p->tensors[vs[si]->tensor];
            if (ts.h != Hi || ts.w != Wi) RY_FAIL("conv: concatenated inputs must share one pixel grid");
            const cuuint64_t cs = (cuuint64_t)ts.d.channels;
            cuuint64_t dims[4] = {(cuuint64_t)(cp.n_src > 1 ? vs[si]->c_len : d.cin), P, 1, 1};
            cuuint64_t str[3] = {cs * esz, P * cs * esz, P * cs * esz};
            if (a.pool) {
                dims[1] = (cuuint64_t)Wi; dims[2] = (cuuint64_t)Hi; dims[3] = (cuuint64_t)B;
                str[1] = (cuuint64_t)Wi * cs * esz; str[2] = (cuuint64_t)Hi * Wi * cs * esz;
            }
            if (encode_map(&m, bf(p, vs[si]->tensor) + vs[si]->c_off, 4, dims, str, box, cp.kb, vs[si]->c_len == ts.d.channels)) return 1;
            maps.push_back(m);
            if (cp.n_src > 1) {
                const int nb = (cp.src_len[si] + cp.kb - 1) / cp.kb;
                for (int j = 0; j < nb; ++j, ++blk) {
                    a.kb_map[blk] = (int8_t)si;
                    a.kb_coord[blk] = (int16_t)(j * cp.kb);
                    a.kb_ks[blk] = (int8_t)((std::min(cp.kb, cp.src_len[si] - j * cp.kb) + 15) / 16);
                }
            }
        }
        a.tap_map[0] = 0; a.tap_dh[0] = 0; a.tap_dw[0] = 0;
    } else if (s == 1 && !no_halo && Wo % kHaloTw == 0 &&
               (Ho % kHaloTh == 0 || (d.cout <= 128 && Ho * 5 >= cdiv(Ho, kHaloTh) * kHaloTh * 4))) {
        // (partial last tile row: worth it for the small-N layers that are bound by L2 re-reads, not by the tensor pipe)
        // halo mode: one (8+2) x (16+2) pixel box per K block, the nine taps are descriptor windows into it
        a.a_mode = A_HALO;
        a.tw = kHaloTw; a.th = kHaloTh; a.tn = 1;
        a.halo_w = kHaloTw + 2;
        a.Wo = Wo; a.Ho = Ho; a.Bo = B;
        const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)B};
        const cuuint64_t str[3] = {ctot * esz, (cuuint64_t)Wi * ctot * esz, (cuuint64_t)Hi * Wi * ctot * esz};
        const cuuint32_t box[4] = {(cuuint32_t)cp.kb, (cuuint32_t)(kHaloTw + 2), (cuuint32_t)(kHaloTh + 2), 1};
        if (encode_map(&m, in_base, 4, dims, str, box, cp.kb, in_dense)) return 1;
        maps.push_back(m);
    } else {
        pick_tile(B, Ho, Wo, &a.tw, &a.th, &a.tn);
        a.Wo = Wo; a.Ho = Ho; a.Bo = B;
        const cuuint32_t box[4] = {(cuuint32_t)cp.kb, (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
        if (s == 1) {
            const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)B};
            const cuuint64_t str[3] = {ctot * esz, (cuuint64_t)Wi * ctot * esz, (cuuint64_t)Hi * Wi * ctot * esz};
            if (encode_map(&m, in_base, 4, dims, str, box, cp.kb, in_dense)) return 1;
            maps.push_back(m);
            for (int t = 0; t < 9; ++t) { a.tap_map[t] = 0; a.tap_dh[t] = (int8_t)(t / 3 - 1); a.tap_dw[t] = (int8_t)(t % 3 - 1); }
        } else {
            // stride 2: the four parity phases of the input are unit-stride tensors with doubled pitches
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw) {
                    const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)((Wi - pw + 1) / 2), (cuuint64_t)((Hi - ph + 1) / 2),
                                                (cuuint64_t)B};
                    const cuuint64_t str[3] = {2 * ctot * esz, 2 * (cuuint64_t)Wi * ctot * esz, (cuuint64_t)Hi * Wi * ctot * esz};
                    if (encode_map(&m, in_base + ((size_t)ph * Wi + pw) * ctot, 4, dims, str, box, cp.kb, false)) return 1;
                    maps.push_back(m);
                }
            for (int t = 0; t < 9; ++t) {
                const int kh = t / 3, kw = t % 3;
                const int ph = kh == 1 ? 0 : 1, pw = kw == 1 ? 0 : 1;
                a.tap_map[t] = (int8_t)(ph * 2 + pw);
                a.tap_dh[t] = (int8_t)(kh == 0 ? -1 : 0);
                a.tap_dw[t] = (int8_t)(kw == 0 ? -1 : 0);
            }
        }
    }
    op.n_amaps = (int)maps.size() - op.tmap_first;
    a.n_acc = (a.a_mode == A_HALO && 4 * cp.BN <= 512) ? 4 : 2;
    if (acc_env && atoi(acc_env) == 4 && 4 * cp.BN <= 512) a.n_acc = 4;
    if (acc_env && atoi(acc_env) == 2) a.n_acc = 2;
    a.tiles_w = cdiv(a.Wo, a.tw); a.tiles_h = cdiv(a.Ho, a.th); a.tiles_n = cdiv(a.Bo, a.tn);
    a.div_hw = (uint64_t)(((1ull << 40) + (uint64_t)a.img_hw - 1) / (uint64_t)a.img_hw);
    a.div_imgw = a.img_w <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)a.img_w - 1) / (uint64_t)a.img_w);
    {
        auto magic = [](int d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
        a.div_nt = magic(a.n_ntiles); a.div_tw = magic(a.tiles_w); a.div_th = magic(a.tiles_h);
        if ((long)a.tiles_w * a.tiles_h * a.tiles_n * a.n_ntiles >= (1L << 20)) RY_FAIL("conv: too many tiles for the fast tile decode");
    }
    {   // weights: [cout_pad][k_pad]
        const cuuint64_t dims[2] = {(cuuint64_t)cp.k_pad, (cuuint64_t)cp.cout_pad};
        const cuuint64_t str[1] = {(cuuint64_t)cp.k_pad * esz};
        const cuuint32_t box[2] = {(cuuint32_t)cp.kb, (cuuint32_t)cp.BN};
        if (encode_map(&m, p->d_weights + cp.w_dev, 2, dims, str, box, cp.kb)) return 1;
        maps.push_back(m);
    }
    a.bias = wf(p, cp.b_dev);
    a.cout = d.cout;
    a.act = d.act;
    const long tiles = (long)a.tiles_w * a.tiles_h * a.tiles_n * a.n_ntiles;
    op.grid = (int)std::min<long>(tiles, num_sms());
    op.launches = 1;
    if (d.kind == RY_OP_DETECT) {
        a.no = p->nc + 5;
        a.na = d.cout / a.no;
        a.det_stride = d.fparam[0];
        for (int i = 0; i < 6; ++i) a.anchors[i] = d.fparam[1 + i];
        if (cp.BN != 32) RY_FAIL("detect: na*(nc+5) must be in (16, 32] for the fused decode epilogue");
        a.n_groups = 2;
        if (conv_plan_smem(a, 8)) RY_FAIL("detect: shared memory plan failed");
        return 0;
    }
    const Tensor &tout = p->tensors[d.out0.tensor];
    if (tout.h != (a.pool ? Ho / 2 : Ho) || tout.w != (a.pool ? Wo / 2 : Wo)) RY_FAIL("conv: output tensor level does not match stride");
    if (a.pool && d.out1.tensor >= 0) RY_FAIL("conv: fused max-pool with a split store is not built");
    if (d.cout % 8 != 0) RY_FAIL("conv: cout must be a multiple of 8");
    OutPiece pieces[2];
    int n_pieces = 1;
    cuuint64_t c_end = (cuuint64_t)tout.d.channels;
    if (d.out1.tensor >= 0) {
        const Tensor &t1 = p->tensors[d.out1.tensor];
        if (cp.n_ntiles != 1 || d.out0.c_len + d.out1.c_len != d.cout || t1.h != Ho || t1.w != Wo) RY_FAIL("conv: bad split store");
        pieces[0] = {0, d.out0.c_len, d.out0.c_off, 0};
        pieces[1] = {d.out0.c_len, d.out1.c_len, d.out1.c_off, 1};   // may be another tensor (two 1x1 convs of one input merged)
        n_pieces = 2;
    } else if (cp.n_ntiles == 1) {
        pieces[0] = {0, d.cout, d.out0.c_off, 0};
    } else {
        pieces[0] = {0, cp.BN, d.out0.c_off, 0};                    // per N tile; channels past the view are clipped by the map
        c_end = (cuuint64_t)(d.out0.c_off + d.out0.c_len);
    }
    a.bv_bytes = (cp.n_src <= 1 && d.in2.tensor >= 0) ? 4 * 2 * cp.cout_pad * 4 : 0;
    a.n_img = B;
    a.tile_contig = a.bv_bytes ? 1 : 0;
    {
        static const bool no_pin = getenv("RY_CONV_NO_PIN") != nullptr;
        a.n_pinned = (!no_pin && a.n_ntiles > 1 && !a.tile_contig && op.grid % a.n_ntiles == 0) ? 1 : 0;
    }
    int widths[8], n_widths = 0;
    int max_cols = 64;
    build_segments(a, pieces, n_pieces, max_cols, widths, &n_widths);
    auto max_w = [&]() { int m = 8; for (int i = 0; i < n_widths; ++i) m = std::max(m, widths[i] % 1000); return m; };
    if (conv_plan_smem(a, max_w())) RY_FAIL("conv: shared memory plan failed");
    if (!a.ep_teams && !a.b_resident && a.a_mode == A_BOX && a.a_stages < 4 && max_w() > 32) {   // big tiles: trade staging for pipeline depth
        build_segments(a, pieces, n_pieces, 32, widths, &n_widths);
        if (conv_plan_smem(a, max_w())) RY_FAIL("conv: shared memory plan failed");
    }
    if (a.nseg[0] > kConvMaxSegs || a.nseg[1] > kConvMaxSegs) RY_FAIL("conv: too many store segments");
    {
        static const char *g_env = getenv("RY_CONV_GROUPS");
        a.n_groups = (a.ep_teams && 4 * cp.BN <= 512) ? 4 : 2;
        if (g_env && atoi(g_env) == 2) a.n_groups = 2;
        if (a.n_groups == 4) a.n_acc = 4;
    }
    for (int wi = 0; wi < n_widths; ++wi) {
        const int slot = widths[wi] / 1000, wd = widths[wi] % 1000;
        const int ot = slot ? d.out1.tensor : d.out0.tensor;
        const cuuint64_t oc = (cuuint64_t)p->tensors[ot].d.channels;
        const cuuint64_t ce = n_pieces == 2 ? oc : c_end;
        if (a.pool) {
            const cuuint64_t dims[4] = {ce, (cuuint64_t)(Wo / 2), (cuuint64_t)(Ho / 2), (cuuint64_t)B};
            const cuuint64_t str[3] = {oc * esz, (cuuint64_t)(Wo / 2) * oc * esz, (cuuint64_t)(Ho / 2) * (Wo / 2) * oc * esz};
            const cuuint32_t box[4] = {(cuuint32_t)wd, (cuuint32_t)(a.tw / 2), (cuuint32_t)(a.th / 2), (cuuint32_t)a.tn};
            if (encode_map(&m, bf(p, ot), 4, dims, str, box, wd)) return 1;
        } else if (k == 1) {
            const cuuint64_t P = (cuuint64_t)a.Wo;
            const cuuint64_t dims[4] = {ce, P, 1, 1};
            const cuuint64_t str[3] = {oc * esz, P * oc * esz, P * oc * esz};
            const cuuint32_t box[4] = {(cuuint32_t)wd, 128, 1, 1};
            if (encode_map(&m, bf(p, ot), 4, dims, str, box, wd)) return 1;
        } else {
            const cuuint64_t dims[4] = {ce, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B};
            const cuuint64_t str[3] = {oc * esz, (cuuint64_t)Wo * oc * esz, (cuuint64_t)Ho * Wo * oc * esz};
            const cuuint32_t box[4] = {(cuuint32_t)wd, (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
            if (encode_map(&m, bf(p, ot), 4, dims, str, box, wd)) return 1;
        }
        maps.push_back(m);
    }
    if (cp.n_src <= 1 && d.in1.tensor >= 0 && d.in2.tensor >= 0) RY_FAIL("conv: residual and per-image vector in one epilogue is not built");
    if (cp.n_src <= 1 && d.in1.tensor >= 0) {
        a.res = bf(p, d.in1.tensor);
        a.res_cs = p->tensors[d.in1.tensor].d.channels;
        a.res_off = d.in1.c_off;
    }
    if (cp.n_src <= 1 && d.in2.tensor >= 0) {
        a.bvec = f32(p, d.in2.tensor);
        a.bvec_cs = p->tensors[d.in2.tensor].d.channels;
        a.bvec_off = d.in2.c_off;
    }
    return 0;
}

// Shape-dependent part of a CONV_CHAIN op.
int bind_chain(ry_plan *p, Op &op, std::vector<CUtensorMap> &maps) {
    const ry_op_desc &d = op.d;
    const ConvPacked &cp = op.cp;
    const Tensor &tin = p->tensors[d.in0.tensor];
    const int H = tin.h, W = tin.w, B = p->B;
    if (W % kHaloTw != 0) RY_FAIL("conv chain: map width must be a multiple of 8");
    if (cp.cblk != 1 || cp.n_ntiles != 1) RY_FAIL("conv chain: the 3x3 stage needs cin, cout <= 64");
    ChainArgs &a = op.ch;
    memset(&a, 0, sizeof(a));
    a.n_stages = 1 + d.n_post;
    a.kb = cp.kb;
    a.halo_w = kHaloTw + 2;
    a.tiles_w = W / kHaloTw; a.tiles_h = cdiv(H, kHaloTh); a.tiles_n = B;
    auto magic = [](int dd) -> uint32_t { return dd <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)dd - 1) / (uint64_t)dd); };
    a.div_tw = magic(a.tiles_w); a.div_th = magic(a.tiles_h);
    if ((long)a.tiles_w * a.tiles_h * a.tiles_n >= (1L << 20)) RY_FAIL("conv chain: too many tiles");
    const ry_view *outs[3] = {&d.out0, &d.out1, &d.out2};
    int prev = d.cout;
    for (int s = 0; s < a.n_stages; ++s) {
        ChainStage &st = a.stage[s];
        st.ncol = s == 0 ? d.cout : d.post_cout[s - 1];
        st.N = s == 0 ? cp.BN : op.post_n[s - 1];
        st.act = s == 0 ? d.act : d.post_act[s - 1];
        st.store = outs[s]->tensor >= 0 ? 1 : 0;
        st.chan = outs[s]->c_off;
        st.swz = st.ncol == 64 ? 7 : (st.ncol == 32 ? 3 : (st.ncol == 16 ? 1 : 0));
        st.kb_next = s + 1 < a.n_stages ? op.post_kb[s] : 0;
        st.ks = s == 0 ? cp.ksteps_last : (prev + 15) / 16;
        if (s > 0) {
            st.w_bytes = op.post_wbytes[s - 1];
            st.w_img = p->d_weights + op.post_w_dev[s - 1];
            st.bias = wf(p, op.post_b_dev[s - 1]);
        } else {
            st.bias = wf(p, cp.b_dev);
        }
        if (st.store && outs[s]->c_len != st.ncol) RY_FAIL("conv chain: store view width does not match the stage");
        prev = st.ncol;
    }
    if (chain_plan_smem(a)) RY_FAIL("conv chain: shared memory / TMEM plan failed");
    op.tmap_first = (int)maps.size();
    const size_t esz = 2;
    CUtensorMap m;
    {
        const cuuint64_t ctot = (cuuint64_t)tin.d.channels;
        const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t str[3] = {ctot * esz, (cuuint64_t)W * ctot * esz, (cuuint64_t)H * W * ctot * esz};
        const cuuint32_t box[4] = {(cuuint32_t)cp.kb, (cuuint32_t)(kHaloTw + 2), (cuuint32_t)(kHaloTh + 2), 1};
        if (encode_map(&m, bf(p, d.in0.tensor) + d.in0.c_off, 4, dims, str, box, cp.kb, d.in0.c_len == tin.d.channels)) return 1;
        maps.push_back(m);
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)cp.k_pad, (cuuint64_t)cp.cout_pad};
        const cuuint64_t str[1] = {(cuuint64_t)cp.k_pad * esz};
        const cuuint32_t box[2] = {(cuuint32_t)cp.kb, (cuuint32_t)cp.BN};
        if (encode_map(&m, p->d_weights + cp.w_dev, 2, dims, str, box, cp.kb)) return 1;
        maps.push_back(m);
    }
    for (int s = 0; s < a.n_stages; ++s) {          // one map slot per stage (unused slots repeat the weight map)
        if (a.stage[s].store) {
            const Tensor &to = p->tensors[outs[s]->tensor];
            if (to.h != H || to.w != W) RY_FAIL("conv chain: output tensor level mismatch");
            const cuuint64_t oc = (cuuint64_t)to.d.channels;
            const cuuint64_t dims[4] = {oc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
            const cuuint64_t str[3] = {oc * esz, (cuuint64_t)W * oc * esz, (cuuint64_t)H * W * oc * esz};
            const cuuint32_t box[4] = {(cuuint32_t)a.stage[s].ncol, (cuuint32_t)kHaloTw, (cuuint32_t)kHaloTh, 1};
            if (encode_map(&m, bf(p, outs[s]->tensor), 4, dims, str, box, a.stage[s].ncol)) return 1;
        }
        maps.push_back(m);
    }
    const long tiles = (long)a.tiles_w * a.tiles_h * a.tiles_n;
    op.grid = (int)std::min<long>(tiles, num_sms());
    op.launches = 1;
    return 0;
}

int run_op(ry_plan *p, int idx, int last, const float *image, float *pred, float *raws[3], cudaStream_t st) {
    Op &op = p->ops[idx];
    const ry_op_desc &d = op.d;
    const int B = p->B;
    switch (d.kind) {
        case RY_OP_STEM: {
            if (!image) RY_FAIL("stem: image pointer is NULL");
            if (reinterpret_cast<uintptr_t>(image) & (p->image_u8 ? 3 : 15)) RY_FAIL("stem: the image must be 16-byte aligned (4-byte for uint8)");
            const Tensor &to = p->tensors[d.out0.tensor];
            if (stem_launch(image, p->image_u8, wf(p, op.dev[0]), wf(p, op.dev[1]), bf(p, d.out0.tensor), to.d.channels, d.out0.c_off, d.cout, B,
                            p->H, p->W, st))
                RY_FAIL("stem: unsupported cout");
            break;
        }
        case RY_OP_CONV: {
            ConvArgs a = op.ca;
            a.amap = p->d_tmaps + op.tmap_first;
            a.wmap = a.amap + op.n_amaps;
            a.omap = a.wmap + 1;
            conv_launch(a, op.grid, st);
            break;
        }
        case RY_OP_CONV_CHAIN: {
            ChainArgs a = op.ch;
            a.amap = p->d_tmaps + op.tmap_first;
            a.wmap = a.amap + 1;
            a.omap = a.amap + 2;
            chain_launch(a, op.grid, st);
            break;
        }
        case RY_OP_DETECT: {
            if (!pred) RY_FAIL("detect: pred pointer is NULL");
            ConvArgs a = op.ca;
            a.amap = p->d_tmaps + op.tmap_first;
            a.wmap = a.amap + op.n_amaps;
            a.omap = a.wmap;
            a.pred = pred;
            a.raw = raws[d.level_idx];
            a.cand_mask = p->filter_mask;                        // ry_decode_filter: obj > conf ballots of the decoded candidates
            a.cand_conf = p->filter_conf;
            a.mask_words = (p->n_cand + 31) / 32;
            conv_launch(a, op.grid, st);
            break;
        }
        case RY_OP_DW5: {
            const Tensor &ti = p->tensors[d.in0.tensor], &to = p->tensors[d.out0.tensor];
            const int C = d.cin;
            const bool split = d.in1.tensor >= 0;
            const int half = split ? d.in0.c_len : C;
            dw5_launch(bf(p, d.in0.tensor), ti.d.channels, d.in0.c_off, split ? d.in1.c_off : 0, bf(p, d.out0.tensor),
                       to.d.channels, d.out0.c_off, split ? d.out1.c_off : 0, wf(p, op.dev[0]), wf(p, op.dev[1]), C, half, B,
                       ti.h, ti.w, d.act, op.tmap_first >= 0 ? p->d_tmaps + op.tmap_first : nullptr, st);
            break;
        }
        case RY_OP_MAXPOOL2: {
            const Tensor &ti = p->tensors[d.in0.tensor], &to = p->tensors[d.out0.tensor];
            maxpool2_launch(bf(p, d.in0.tensor), ti.d.channels, d.in0.c_off, bf(p, d.out0.tensor), to.d.channels, d.out0.c_off,
                            d.in0.c_len, B, ti.h, ti.w, st);
            break;
        }
        case RY_OP_SPP: {
            const Tensor &ti = p->tensors[d.in0.tensor], &to = p->tensors[d.out0.tensor];
            spp_launch(bf(p, d.in0.tensor), ti.d.channels, d.in0.c_off, bf(p, d.out0.tensor), to.d.channels, d.out0.c_off,
                       d.out1.c_off, d.out2.c_off, d.in0.c_len, B, ti.h, ti.w, st);
            break;
        }
        case RY_OP_UPSAMPLE2: {
            const Tensor &ti = p->tensors[d.in0.tensor], &to = p->tensors[d.out0.tensor];
            upsample2_launch(bf(p, d.in0.tensor), ti.d.channels, d.in0.c_off, bf(p, d.out0.tensor), to.d.channels, d.out0.c_off,
                             d.in0.c_len, B, ti.h, ti.w, st);
            break;
        }
        case RY_OP_CA: {
            const Tensor &ti = p->tensors[d.in0.tensor], &to = p->tensors[d.out0.tensor];
            ca_launch(bf(p, d.in0.tensor), ti.d.channels, d.in0.c_off, f32(p, d.out0.tensor), to.d.channels, d.out0.c_off,
                      wf(p, op.dev[0]), wf(p, op.dev[1]), d.in0.c_len, B, ti.h * ti.w, reinterpret_cast<float *>(p->ws + p->scratch_off), st);
            break;
        }
        case RY_OP_ATTN_QK: {
            const Tensor &ti = p->tensors[d.in0.tensor];
            attn_qk_launch(bf(p, d.in0.tensor), ti.d.channels, d.in0.c_off, d.cin, d.cout, (size_t)B * ti.h * ti.w, wf(p, op.dev[0]),
                           wf(p, op.dev[1]), wf(p, op.dev[2]), wf(p, op.dev[3]), wf(p, op.dev[4]), wf(p, op.dev[5]),
                           f32(p, d.out0.tensor), f32(p, d.out1.tensor), st);
            break;
        }
        case RY_OP_CRISSCROSS:
        case RY_OP_VERTICAL: {
            auto params = [&](const Op &o) {
                const ry_op_desc &od = o.d;
                const Tensor &ti = p->tensors[od.in0.tensor], &to = p->tensors[od.out0.tensor];
                AttnParams ap;
                ap.x = bf(p, od.in0.tensor); ap.x_cs = ti.d.channels; ap.x_off = od.in0.c_off;
                ap.C = od.cin; ap.Cq = od.cin / 8; ap.B = B; ap.H = ti.h; ap.W = ti.w;
                ap.qk = wf(p, o.dev[4]);
                ap.wv = wf(p, o.dev[0]); ap.bv = wf(p, o.dev[1]); ap.s1 = wf(p, o.dev[2]); ap.t1 = wf(p, o.dev[3]);
                ap.gamma = od.fparam[0];
                ap.out = bf(p, od.out0.tensor); ap.out_cs = to.d.channels; ap.out_off = od.out0.c_off;
                ap.scratch = reinterpret_cast<float *>(p->ws + p->scratch_off);
                return ap;
            };
            const AttnParams ap = params(op);
            int rc;
            if (d.kind == RY_OP_CRISSCROSS) {
                // CCVA = m1(m(x)): when the next op of this run is the VerticalAttention reading this output, its energy pass is
                // fused into the column pass (the vertical op then only runs its value pass)
                const Op *nx = (idx + 1 < last) ? &p->ops[idx + 1] : nullptr;
                const bool feeds = nx && nx->d.kind == RY_OP_VERTICAL && nx->d.in0.tensor == d.out0.tensor && nx->d.in0.c_off == d.out0.c_off &&
                                   nx->d.in0.c_len == d.out0.c_len && nx->d.cin == d.cin;
                int done = 0;
                if (feeds) {
                    const AttnParams nxp = params(*nx);
                    rc = crisscross_launch(ap, &nxp, &done, st);
                } else {
                    rc = crisscross_launch(ap, nullptr, nullptr, st);
                }
                p->energies_ready_op = done ? idx + 1 : -1;
            } else {
                rc = vertical_launch(ap, p->energies_ready_op == idx ? 1 : 0, st);
                p->energies_ready_op = -1;
            }
            if (rc) RY_FAIL("attention: unsupported shape (shared memory)");
            break;
        }
        default: RY_FAIL("unknown op kind");
    }
    return 0;
}

size_t tensor_bytes(const Tensor &t, int B, int H, int W) {
    const size_t e = t.d.dtype == RY_BF16 ? 2 : 4;
    if (t.d.kind == RY_T_MAP) return (size_t)B * (H >> t.d.level) * (W >> t.d.level) * t.d.channels * e;
    if (t.d.kind == RY_T_VEC) return (size_t)B * t.d.channels * e;
    return 0;
}

size_t scratch_bytes(const ry_plan *p, int B, int H, int W) {
    size_t m = 0;
    for (const Op &op : p->ops)
        if (op.d.kind == RY_OP_CRISSCROSS || op.d.kind == RY_OP_VERTICAL) {
            const Tensor &t = p->tensors[op.d.in0.tensor];
            m = std::max(m, attn_scratch_bytes(B, H >> t.d.level, W >> t.d.level, op.d.cin));
        } else if (op.d.kind == RY_OP_CA) {
            m = std::max(m, ca_scratch_bytes(B, op.d.cin));
        }
    return m;
}

}  // namespace
}  // namespace ry

using namespace ry;

extern "C" {

int ry_abi_version(void) { return RY_ABI_VERSION; }
int ry_abi_sizeof(int which) { return which == 0 ? (int)sizeof(ry_tensor_desc) : (which == 1 ? (int)sizeof(ry_op_desc) : -1); }
const char *ry_last_error(void) { return g_error.c_str(); }

int ry_plan_create(const ry_tensor_desc *tensors, int n_tensors, const ry_op_desc *ops, int n_ops, const void *weights_host,
                   size_t weight_bytes, int nc, int device, ry_plan **out) {
    if (!tensors || !ops || !out || n_tensors <= 0 || n_ops <= 0) RY_FAIL("ry_plan_create: bad arguments");
    RY_CUDA(cudaSetDevice(device));
    ry_plan *p = new ry_plan();
    p->device = device;
    p->nc = nc;
    p->tensors.resize(n_tensors);
    for (int i = 0; i < n_tensors; ++i) p->tensors[i].d = tensors[i];
    p->ops.resize(n_ops);
    const unsigned char *host = static_cast<const unsigned char *>(weights_host);
    Blob blob;
    int rc = 0;
    for (int i = 0; i < n_ops && !rc; ++i) {
        Op &op = p->ops[i];
        op.d = ops[i];
        const ry_op_desc &d = op.d;
        const ry_view *vs[6] = {&d.in0, &d.in1, &d.in2, &d.out0, &d.out1, &d.out2};
        for (int v = 0; v < 6; ++v)
            if (view_ok(p, *vs[v], v == 0 || (v == 3 && d.kind != RY_OP_CONV_CHAIN))) { set_error("op " + std::to_string(i) + ": bad tensor view"); rc = 1; }
        if (rc) break;
        switch (d.kind) {
            case RY_OP_CONV:
            case RY_OP_DETECT: rc = pack_conv(d, host, weight_bytes, blob, op.cp); break;
            case RY_OP_CONV_CHAIN: {
                if (d.ksize != 3 || d.stride != 1 || d.n_post < 1 || d.n_post > 2) { set_error("conv chain: 3x3 s1 main conv with 1 or 2 fused 1x1 stages"); rc = 1; break; }
                rc = pack_conv(d, host, weight_bytes, blob, op.cp);
                int prev = d.cout;
                for (int i = 0; i < d.n_post && !rc; ++i) {
                    rc = pack_post(host, weight_bytes, d.aux_off[2 * i], d.aux_off[2 * i + 1], prev, d.post_cout[i], blob, op, i);
                    prev = d.post_cout[i];
                }
                break;
            }
            case RY_OP_STEM: {
                if (d.cin != 3 || d.ksize != 3 || d.stride != 2) { set_error("stem: expects 3x3 s2 on 3 channels"); rc = 1; break; }
                if (d.w_off < 0 || (size_t)d.w_off + (size_t)d.cout * 27 * 4 > weight_bytes) { set_error("stem: weight offset"); rc = 1; break; }
                const float *w = reinterpret_cast<const float *>(host + d.w_off);
                std::vector<float> w27((size_t)27 * d.cout);
                for (int co = 0; co < d.cout; ++co)
                    for (int t = 0; t < 27; ++t) w27[(size_t)t * d.cout + co] = w[(size_t)co * 27 + t];
                op.dev[0] = blob.add(w27.data(), w27.size() * 4);
                rc = copy_f32(host, weight_bytes, d.b_off, d.cout, blob, &op.dev[1]);
                break;
            }
            case RY_OP_DW5: {
                const int C = d.cin;
                if (d.w_off < 0 || (size_t)d.w_off + (size_t)C * 25 * 4 > weight_bytes) { set_error("dw5: weight offset"); rc = 1; break; }
                const float *w = reinterpret_cast<const float *>(host + d.w_off);
                std::vector<float> wt((size_t)25 * C);
                for (int c = 0; c < C; ++c)
                    for (int t = 0; t < 25; ++t) wt[(size_t)t * C + c] = w[(size_t)c * 25 + t];
                op.dev[0] = blob.add(wt.data(), wt.size() * 4);
                rc = copy_f32(host, weight_bytes, d.b_off, C, blob, &op.dev[1]);
                break;
            }
            case RY_OP_CA: {
                const int C = d.cin;
                rc = copy_f32(host, weight_bytes, d.w_off, (size_t)(C / 16) * C, blob, &op.dev[0]) ||
                     copy_f32(host, weight_bytes, d.aux_off[0], (size_t)C * (C / 16), blob, &op.dev[1]);
                break;
            }
            case RY_OP_ATTN_QK: {
                const int Cq = d.cout;
                rc = copy_f32(host, weight_bytes, d.w_off, (size_t)Cq * 8, blob, &op.dev[0]) ||
                     copy_f32(host, weight_bytes, d.b_off, Cq, blob, &op.dev[1]) ||
                     copy_f32(host, weight_bytes, d.aux_off[0], (size_t)Cq * 8, blob, &op.dev[2]) ||
                     copy_f32(host, weight_bytes, d.aux_off[1], Cq, blob, &op.dev[3]) ||
                     copy_f32(host, weight_bytes, d.aux_off[2], Cq, blob, &op.dev[4]) ||
                     copy_f32(host, weight_bytes, d.aux_off[3], Cq, blob, &op.dev[5]);
                break;
            }
            case RY_OP_CRISSCROSS:
            case RY_OP_VERTICAL: {
                const int C = d.cin;
                rc = copy_f32(host, weight_bytes, d.w_off, C, blob, &op.dev[0]) || copy_f32(host, weight_bytes, d.b_off, C, blob, &op.dev[1]) ||
                     copy_f32(host, weight_bytes, d.aux_off[0], C, blob, &op.dev[2]) ||
                     copy_f32(host, weight_bytes, d.aux_off[1], C, blob, &op.dev[3]) ||
                     copy_f32(host, weight_bytes, d.aux_off[2], (size_t)(C / 8) * 20, blob, &op.dev[4]);   // packed q/k conv parameters
                break;
            }
            default: break;
        }
    }
    if (rc) { delete p; return 1; }
    p->weight_bytes = std::max<size_t>(blob.data.size(), 256);
    if (cudaMalloc(&p->d_weights, p->weight_bytes) != cudaSuccess) { delete p; RY_FAIL("cudaMalloc of the weight buffer failed"); }
    if (cudaMemcpy(p->d_weights, blob.data.data(), blob.data.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(p->d_weights);
        delete p;
        RY_FAIL("weight upload failed");
    }
    *out = p;
    return 0;
}

void ry_plan_destroy(ry_plan *p) {
    if (!p) return;
    if (p->d_weights) cudaFree(p->d_weights);
    if (p->d_tmaps) cudaFree(p->d_tmaps);
    for (auto &e : p->events) cudaEventDestroy(e);
    delete p;
}

int ry_plan_workspace_bytes(ry_plan *p, int B, int H, int W, size_t *bytes) {
    if (!p || !bytes || B <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32) RY_FAIL("workspace_bytes: H and W must be positive multiples of 32");
    size_t total = 0;
    for (const Tensor &t : p->tensors) total += align_up(tensor_bytes(t, B, H, W), 1024);
    total += align_up(scratch_bytes(p, B, H, W), 1024);
    *bytes = total + 1024;
    return 0;
}

int ry_plan_bind(ry_plan *p, int B, int H, int W, void *workspace, size_t workspace_bytes) {
    size_t need = 0;
    if (ry_plan_workspace_bytes(p, B, H, W, &need)) return 1;
    if (!workspace || workspace_bytes < need) RY_FAIL("bind: workspace too small");
    RY_CUDA(cudaSetDevice(p->device));
    p->B = B; p->H = H; p->W = W;
    unsigned char *base = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(workspace), 1024));
    p->ws = base;
    p->ws_bytes = workspace_bytes - (size_t)(base - static_cast<unsigned char *>(workspace));
    size_t off = 0;
    for (Tensor &t : p->tensors) {
        t.h = t.d.kind == RY_T_MAP ? (H >> t.d.level) : 1;
        t.w = t.d.kind == RY_T_MAP ? (W >> t.d.level) : 1;
        t.bytes = tensor_bytes(t, B, H, W);
        t.offset = off;
        off += align_up(t.bytes, 1024);
    }
    p->scratch_off = off;
    std::vector<CUtensorMap> maps;
    p->n_cand = 0;
    for (Op &op : p->ops) {
        op.launches = 1;
        if (op.d.kind == RY_OP_CONV || op.d.kind == RY_OP_DETECT) {
            if (bind_conv(p, op, maps)) return 1;
        }
        if (op.d.kind == RY_OP_CONV_CHAIN && bind_chain(p, op, maps)) return 1;
        if (op.d.kind == RY_OP_DW5) {
            // input tensor as {C, W, H, B}; one box = 8 channels x the halo of a tile (dw5_mma_kernel)
            const Tensor &ti = p->tensors[op.d.in0.tensor];
            const cuuint64_t cs = (cuuint64_t)ti.d.channels;
            const cuuint64_t dims[4] = {cs, (cuuint64_t)ti.w, (cuuint64_t)ti.h, (cuuint64_t)B};
            const cuuint64_t str[3] = {cs * 2, (cuuint64_t)ti.w * cs * 2, (cuuint64_t)ti.h * ti.w * cs * 2};
            const cuuint32_t box[4] = {8, (cuuint32_t)(dw5_tile_w(ti.w) + 4), (cuuint32_t)dw5_halo_rows(), 1};
            CUtensorMap m;
            op.tmap_first = (int)maps.size();
            if (encode_map(&m, bf(p, op.d.in0.tensor), 4, dims, str, box, 8, false)) return 1;
            maps.push_back(m);
        }
        if (op.d.kind == RY_OP_CA) op.launches = 2;                                               // partial sums + finish
        if (op.d.kind == RY_OP_CRISSCROSS || op.d.kind == RY_OP_VERTICAL) op.launches = 2;         // two line passes
    }
    for (size_t i = 1; i < p->ops.size(); ++i)                                                    // vertical energy pass fused into the criss-cross column pass
        if (p->ops[i].d.kind == RY_OP_VERTICAL && p->ops[i - 1].d.kind == RY_OP_CRISSCROSS &&
            p->ops[i].d.in0.tensor == p->ops[i - 1].d.out0.tensor && p->ops[i].d.in0.c_off == p->ops[i - 1].d.out0.c_off)
            p->ops[i].launches = 1;
    // Detect rows: level -> anchor -> y -> x (models/yolo.py:152, 166)
    int rows = 0;
    for (Op &op : p->ops)
        if (op.d.kind == RY_OP_DETECT) {
            op.ca.row_off = rows;
            rows += op.ca.na * op.ca.img_hw;
        }
    for (Op &op : p->ops)
        if (op.d.kind == RY_OP_DETECT) op.ca.rows_total = rows;
    p->n_cand = rows;
    if (maps.size() > p->tmaps_cap) {
        if (p->d_tmaps) cudaFree(p->d_tmaps);
        p->d_tmaps = nullptr;
        RY_CUDA(cudaMalloc(&p->d_tmaps, maps.size() * sizeof(CUtensorMap)));
        p->tmaps_cap = maps.size();
    }
    if (!maps.empty()) RY_CUDA(cudaMemcpy(p->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    return 0;
}

int ry_plan_tensor_info(ry_plan *p, int t, size_t *offset, int *h, int *w, int *c, int *dtype) {
    if (!p || t < 0 || t >= (int)p->tensors.size() || !p->ws) RY_FAIL("tensor_info: bad tensor or plan not bound");
    const Tensor &T = p->tensors[t];
    if (offset) *offset = T.offset;
    if (h) *h = T.h;
    if (w) *w = T.w;
    if (c) *c = T.d.channels;
    if (dtype) *dtype = T.d.dtype;
    return 0;
}

int ry_plan_num_candidates(ry_plan *p, int *n) {
    if (!p || !p->ws || !n) RY_FAIL("num_candidates: plan not bound");
    *n = p->n_cand;
    return 0;
}

int ry_plan_launch_count(ry_plan *p, int *n) {
    if (!p || !p->ws || !n) RY_FAIL("launch_count: plan not bound");
    int c = 0;
    for (const Op &op : p->ops) c += op.launches;
    *n = c;
    return 0;
}

int ry_run_ops(ry_plan *p, int first, int last, const float *image, float *pred, float *raw0, float *raw1, float *raw2,
               void *stream) {
    if (!p || !p->ws) RY_FAIL("run_ops: plan not bound");
    if (first < 0 || last > (int)p->ops.size() || first > last) RY_FAIL("run_ops: bad op range");
    float *raws[3] = {raw0, raw1, raw2};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p->energies_ready_op = -1;
    for (int i = first; i < last; ++i) {
        if (p->profiling) cudaEventRecord(p->events[2 * i], st);
        if (run_op(p, i, last, image, pred, raws, st)) return 1;
        if (p->profiling) cudaEventRecord(p->events[2 * i + 1], st);
    }
    RY_CUDA(cudaGetLastError());
    return 0;
}

int ry_plan_set_image_dtype(ry_plan *p, int dtype) {
    if (!p || (dtype != RY_F32 && dtype != RY_U8)) RY_FAIL("set_image_dtype: RY_F32 or RY_U8");
    p->image_u8 = dtype == RY_U8 ? 1 : 0;
    return 0;
}

int ry_plan_set_profiling(ry_plan *p, int on) {
    if (!p) RY_FAIL("set_profiling: NULL plan");
    if (on && p->events.empty()) {
        p->events.resize(2 * p->ops.size());
        for (auto &e : p->events) RY_CUDA(cudaEventCreate(&e));
    }
    p->profiling = on != 0;
    return 0;
}

int ry_plan_op_times(ry_plan *p, float *ms_host, int n) {
    if (!p || !ms_host || n != (int)p->ops.size() || p->events.empty()) RY_FAIL("op_times: profiling was never enabled");
    for (int i = 0; i < n; ++i) {
        ms_host[i] = 0.0f;
        if (cudaEventQuery(p->events[2 * i + 1]) == cudaSuccess) cudaEventElapsedTime(&ms_host[i], p->events[2 * i], p->events[2 * i + 1]);
    }
    cudaGetLastError();
    return 0;
}

int ry_forward(ry_plan *p, const float *image, float *pred, float *raw0, float *raw1, float *raw2, void *stream) {
    if (!p) RY_FAIL("forward: NULL plan");
    return ry_run_ops(p, 0, (int)p->ops.size(), image, pred, raw0, raw1, raw2, stream);
}

int ry_decode_filter(ry_plan *p, const float *image, float conf_thres, float *pred, float *raw0, float *raw1, float *raw2,
                     uint32_t *cand_mask, void *stream) {
    if (!p || !p->ws) RY_FAIL("decode_filter: plan not bound");
    if (!cand_mask) RY_FAIL("decode_filter: NULL candidate mask");
    const size_t words = (size_t)p->B * (size_t)((p->n_cand + 31) / 32);
    RY_CUDA(cudaMemsetAsync(cand_mask, 0, words * 4, static_cast<cudaStream_t>(stream)));
    p->filter_mask = cand_mask;
    p->filter_conf = conf_thres;
    const int rc = ry_run_ops(p, 0, (int)p->ops.size(), image, pred, raw0, raw1, raw2, stream);
    p->filter_mask = nullptr;
    return rc;
}

int ry_nms_workspace_bytes(int B, int N, int nc, int multi_label, size_t *bytes) {
    if (!bytes || B <= 0 || N <= 0 || nc <= 0) RY_FAIL("nms_workspace_bytes: bad arguments");
    *bytes = nms_workspace_bytes(B, N, nc, multi_label);
    return 0;
}

int ry_nms_launch_count(int B, int N, int nc, int multi_label, int *n) {
    if (!n || B <= 0 || N <= 0 || nc <= 0) RY_FAIL("nms_launch_count: bad arguments");
    *n = nms_launch_count(B, N, nc, multi_label);
    return 0;
}

int ry_nms(const float *pred, int B, int N, int nc, float conf_thres, double iou_thres, const int32_t *classes_host,
           int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float *out, int32_t *counts, void *workspace,
           size_t workspace_bytes, void *stream) {
    if (!pred || !out || !counts || !workspace) RY_FAIL("ry_nms: NULL pointer");
    return nms_run(pred, nullptr, B, N, nc, conf_thres, iou_thres, classes_host, n_classes, agnostic, multi_label, max_det, max_nms, out,
                   counts, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ry_nms_filtered(const float *pred, const uint32_t *cand_mask, int B, int N, int nc, float conf_thres, double iou_thres,
                    const int32_t *classes_host, int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float *out,
                    int32_t *counts, void *workspace, size_t workspace_bytes, void *stream) {
    if (!pred || !cand_mask || !out || !counts || !workspace) RY_FAIL("ry_nms_filtered: NULL pointer");
    return nms_run(pred, cand_mask, B, N, nc, conf_thres, iou_thres, classes_host, n_classes, agnostic, multi_label, max_det, max_nms,
                   out, counts, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
