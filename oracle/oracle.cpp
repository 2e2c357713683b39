// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Single-threaded CPU restatement of the FIR / running-extrema hot path of
// JuliaImages/ImageFiltering.jl v0.7.12, exported through the same C ABI as the CUDA library
// (include/b2f.h) but built into a SEPARATE shared object (oracle/libb2f_oracle.so).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it; the product library never links or calls anything in this file.
//
// Julia is not installed in the build image, so the reference itself cannot run here.  Parity of
// this restatement is PINNED by the reference's own literal test goldens (tests/golden/*.json,
// transcribed from reference test/*.jl with file:line citations; tests/test_oracle_goldens.py).
// One boundary stays unpinned at the bit level: FixedPointNumbers' N0f8 -> float conversion
// (third-party, not under /root/reference; restated here as correctly rounded i/255).
//
// What is restated, loop for loop (paths relative to /root/reference):
//   pad once          src/imfilter.jl:331-341, src/border.jl:324-347 (padarray/copydata!),
//                     src/border.jl:564-596 (padindex), :644-645 (modrange)
//   padding algebra   src/border.jl:602-642 (lo/hi/calculate_padding/accumulate_padding),
//                     :657-684 (next_shrink/expand/shrink)
//   scheduler         src/imfilter.jl:367-395 (trivial/single/cascade), :419-457 (_imfilter!),
//                     :1317-1329 (tempbuffer: eltype(out) temporaries of the padded size)
//   validation        src/imfilter.jl:592-617
//   dense loop        src/imfilter.jl:624-669      (J column-major ascending, separate mul and add)
//   1-D loop          src/imfilter.jl:671-739      (post / i / pre nest, taps ascending)
//   accumulator type  src/imfilter.jl:630-632, src/utils.jl:122-133
//   tiled + threads   src/imfilter.jl:398-404,460-542,1298-1312 (b2f_oracle_imfilter_tiled)
//   imgradients       src/specialty.jl:39-53
//   mapwindow min/max src/mapwindow.jl:270-333 (generic, copy_win!), :388-481 (extrema_filter)
//   local extrema     src/extrema.jl:125-164 (findlocalextrema), :94-105 (multiLoG slice), :85-90 (blob_LoG plumbing)
//
// Build: g++ -O3 -mavx2 -ffp-contract=off (NO -ffast-math, NO FMA contraction: Julia emits neither).

#include "../include/b2f.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

struct Box {  // inclusive index ranges per axis; axes >= ndim are 0:0
    int64_t lo[B2F_MAXDIM], hi[B2F_MAXDIM];
    int64_t len(int d) const { return hi[d] - lo[d] + 1; }
    bool empty() const {
        for (int d = 0; d < B2F_MAXDIM; ++d)
            if (hi[d] < lo[d]) return true;
        return false;
    }
};

Box axes_of(const b2f_array *a) {
    Box b;
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        if (d < a->ndim) {
            b.lo[d] = a->origin[d];
            b.hi[d] = a->origin[d] + a->dims[d] - 1;
        } else {
            b.lo[d] = b.hi[d] = 0;
        }
    }
    return b;
}

// Column-major strided view over a typed buffer with arbitrary first indices (an OffsetArray).
template <typename S>
struct View {
    S *p;
    Box ax;
    int64_t st[B2F_MAXDIM];
    void init(S *ptr, const Box &b) {
        p = ptr;
        ax = b;
        int64_t s = 1;
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            st[d] = s;
            s *= std::max<int64_t>(b.len(d), 0);
        }
    }
    int64_t count() const {
        int64_t n = 1;
        for (int d = 0; d < B2F_MAXDIM; ++d) n *= std::max<int64_t>(ax.len(d), 0);
        return n;
    }
    inline int64_t off(int64_t i0, int64_t i1, int64_t i2, int64_t i3) const {
        return (i0 - ax.lo[0]) * st[0] + (i1 - ax.lo[1]) * st[1] + (i2 - ax.lo[2]) * st[2] +
               (i3 - ax.lo[3]) * st[3];
    }
};

bool is_int_dtype(int dt) {
    return dt == B2F_U8 || dt == B2F_I16 || dt == B2F_I32 || dt == B2F_I64 || dt == B2F_U16 ||
           dt == B2F_U32;
}
size_t dtype_size(int dt) {
    switch (dt) {
        case B2F_U8: case B2F_N0F8: return 1;
        case B2F_I16: case B2F_U16: return 2;
        case B2F_I32: case B2F_U32: case B2F_F32: return 4;
        default: return 8;
    }
}
bool int_range(int dt, int64_t &lo, int64_t &hi) {
    switch (dt) {
        case B2F_U8: lo = 0; hi = 255; return true;
        case B2F_I16: lo = -32768; hi = 32767; return true;
        case B2F_U16: lo = 0; hi = 65535; return true;
        case B2F_I32: lo = INT32_MIN; hi = INT32_MAX; return true;
        case B2F_U32: lo = 0; hi = UINT32_MAX; return true;
        case B2F_I64: lo = INT64_MIN; hi = INT64_MAX; return true;
    }
    return false;
}

// ---- element conversion img -> S at pad time (src/border.jl:343: dest[i,I] = img[j,J]) -------
// N0f8 -> float: FixedPointNumbers computes reinterpret(x)/rawone in the target float type.
template <typename S>
int load_as(const void *ptr, int dt, int64_t i, S &v) {
    switch (dt) {
        case B2F_U8: v = (S)((const uint8_t *)ptr)[i]; return 0;
        case B2F_N0F8:
            if (std::is_same<S, float>::value) v = (S)((float)((const uint8_t *)ptr)[i] / 255.0f);
            else if (std::is_same<S, double>::value) v = (S)((double)((const uint8_t *)ptr)[i] / 255.0);
            else return B2F_EARG;
            return 0;
        case B2F_I16: v = (S)((const int16_t *)ptr)[i]; return 0;
        case B2F_U16: v = (S)((const uint16_t *)ptr)[i]; return 0;
        case B2F_I32: v = (S)((const int32_t *)ptr)[i]; return 0;
        case B2F_U32: v = (S)((const uint32_t *)ptr)[i]; return 0;
        case B2F_I64: v = (S)((const int64_t *)ptr)[i]; return 0;
        case B2F_F32: {
            float f = ((const float *)ptr)[i];
            if (std::is_integral<S>::value) {
                if (!(std::floor(f) == f) || std::fabs(f) > 9.0e18f) return B2F_EINEXACT;
            }
            v = (S)f;
            return 0;
        }
        case B2F_F64: {
            double f = ((const double *)ptr)[i];
            if (std::is_integral<S>::value) {
                if (!(std::floor(f) == f) || std::fabs(f) > 9.0e18) return B2F_EINEXACT;
            }
            v = (S)f;
            return 0;
        }
    }
    return B2F_EARG;
}

// ---- padindex (src/border.jl:564-596), 0-based position -> 0-based source position ------------
// `n` = axis length.  Returns false for the reflect/n==1 case where the reference divides by zero.
bool pad_source_index(int style, int64_t i, int64_t n, int64_t &src) {
    auto mod = [](int64_t a, int64_t m) { int64_t r = a % m; return r < 0 ? r + m : r; };
    switch (style) {
        case B2F_REPLICATE: src = i < 0 ? 0 : (i >= n ? n - 1 : i); return true;
        case B2F_CIRCULAR: src = mod(i, n); return true;
        case B2F_SYMMETRIC: {  // index table [0..n-1, n-1..0], period 2n
            int64_t m = mod(i, 2 * n);
            src = m < n ? m : 2 * n - 1 - m;
            return true;
        }
        case B2F_REFLECT: {  // index table [0..n-1, n-2..1], period 2n-2
            if (n < 2) return false;
            int64_t m = mod(i, 2 * n - 2);
            src = m < n ? m : 2 * n - 2 - m;
            return true;
        }
    }
    return false;
}

struct StageInfo {
    const b2f_stage *s;
    int64_t lo[B2F_MAXDIM], hi[B2F_MAXDIM];  // tap index range per axis (0:0 on unused axes)
    bool copy;                               // iscopy (src/imfilter.jl:1252-1255)
};

int analyse_stage(const b2f_stage *s, int ndim, StageInfo &si) {
    si.s = s;
    for (int d = 0; d < B2F_MAXDIM; ++d) si.lo[d] = si.hi[d] = 0;
    if (!s->taps && s->kind != B2F_STAGE_LAPLACIAN) return fail(B2F_EARG, "stage has NULL taps");
    if (s->kind == B2F_STAGE_LAPLACIAN) {
        if (s->ndim != ndim)
            return fail(B2F_EDIM, "Laplacian has %d dims, array has %d", s->ndim, ndim);
        for (int d = 0; d < ndim; ++d)
            if (s->len[d] == 3) { si.lo[d] = -1; si.hi[d] = 1; }
        si.copy = false;
        return 0;
    }
    if (s->kind == B2F_STAGE_1D) {
        if (s->axis < 0 || s->axis >= ndim)
            return fail(B2F_EDIM, "1-D stage axis %d out of range for %d-d array", s->axis, ndim);
        if (s->len[s->axis] < 1) return fail(B2F_EARG, "empty kernel factor");
        si.lo[s->axis] = s->lo[s->axis];
        si.hi[s->axis] = s->lo[s->axis] + s->len[s->axis] - 1;
    } else if (s->kind == B2F_STAGE_DENSE) {
        if (s->ndim != ndim)
            return fail(B2F_EDIM, "dense kernel has %d dims, array has %d", s->ndim, ndim);
        for (int d = 0; d < ndim; ++d) {
            if (s->len[d] < 1) return fail(B2F_EARG, "empty kernel");
            si.lo[d] = s->lo[d];
            si.hi[d] = s->lo[d] + s->len[d] - 1;
        }
    } else {
        return fail(B2F_EARG, "unknown stage kind %d", s->kind);
    }
    bool unit = true;
    for (int d = 0; d < B2F_MAXDIM; ++d) unit = unit && si.lo[d] == 0 && si.hi[d] == 0;
    si.copy = unit && s->taps[0] == 1.0;
    return 0;
}

template <typename T> struct StoreCheck {
    static inline bool conv(double acc, T &v) { v = (T)acc; return true; }
    static inline bool conv(float acc, T &v) { v = (T)acc; return true; }
    static inline bool conv(int64_t acc, T &v) { v = (T)acc; return true; }
};
template <> struct StoreCheck<int64_t> {
    static inline bool conv(double acc, int64_t &v) {
        if (!(std::floor(acc) == acc) || std::fabs(acc) > 9.0e18) return false;
        v = (int64_t)acc;
        return true;
    }
    static inline bool conv(float acc, int64_t &v) { return conv((double)acc, v); }
    static inline bool conv(int64_t acc, int64_t &v) { v = acc; return true; }
};

// One stage, valid region R, src -> dst (both S-typed views).  ACC = accumulator type.
// 1-D: src/imfilter.jl:724-739; dense: :650-669 (generalised to N-d by CartesianIndices order).
template <typename S, typename ACC>
int run_stage(const StageInfo &si, const View<S> &src, View<S> &dst, const Box &R, bool int_out,
              int64_t ilo, int64_t ihi) {
    const b2f_stage *s = si.s;
    bool inexact = false;
    if (s->kind == B2F_STAGE_LAPLACIAN) {
        // src/specialty.jl:3-16: tmp = convert(TT, -n*A[I]); tmp += A[I+J]; tmp += A[I-J] per flagged axis
        int nfl = 0;
        int64_t offs[B2F_MAXDIM];
        for (int d = 0; d < B2F_MAXDIM; ++d)
            if (si.hi[d] == 1) offs[nfl++] = src.st[d];
        const ACC n = (ACC)(2 * nfl);
        for (int64_t i3 = R.lo[3]; i3 <= R.hi[3]; ++i3)
        for (int64_t i2 = R.lo[2]; i2 <= R.hi[2]; ++i2)
        for (int64_t i1 = R.lo[1]; i1 <= R.hi[1]; ++i1)
        for (int64_t i0 = R.lo[0]; i0 <= R.hi[0]; ++i0) {
            const S *q = src.p + src.off(i0, i1, i2, i3);
            ACC tmp = -n * (ACC)q[0];
            for (int t = 0; t < nfl; ++t) { tmp += (ACC)q[offs[t]]; tmp += (ACC)q[-offs[t]]; }
            S v;
            if (!StoreCheck<S>::conv(tmp, v)) inexact = true;
            else if (int_out && ((int64_t)v < ilo || (int64_t)v > ihi)) inexact = true;
            dst.p[dst.off(i0, i1, i2, i3)] = v;
        }
        return inexact ? B2F_EINEXACT : 0;
    }
    if (s->kind == B2F_STAGE_1D) {
        const int ax = s->axis;
        const int64_t L = s->len[ax], klo = s->lo[ax];
        std::vector<ACC> k(L);
        for (int64_t j = 0; j < L; ++j) k[j] = (ACC)s->taps[j];
        // loop nest: post / i / pre with "pre" innermost (contiguous for ax>0)
        int64_t lo[B2F_MAXDIM], hi[B2F_MAXDIM];
        for (int d = 0; d < B2F_MAXDIM; ++d) { lo[d] = R.lo[d]; hi[d] = R.hi[d]; }
        const int64_t sst = src.st[ax];
        for (int64_t i3 = lo[3]; i3 <= hi[3]; ++i3)
        for (int64_t i2 = lo[2]; i2 <= hi[2]; ++i2)
        for (int64_t i1 = lo[1]; i1 <= hi[1]; ++i1) {
            const S *sp = src.p + src.off(lo[0], i1, i2, i3) + klo * sst;
            S *dp = dst.p + dst.off(lo[0], i1, i2, i3);
            const int64_t n0 = hi[0] - lo[0] + 1;
            const int64_t s0 = 1;  // axis-0 stride of both views
            for (int64_t x = 0; x < n0; ++x) {
                ACC tmp = (ACC)0;
                const S *q = sp + x * s0;
                for (int64_t j = 0; j < L; ++j) tmp += (ACC)q[j * sst] * k[j];
                S v;
                if (!StoreCheck<S>::conv(tmp, v)) inexact = true;
                else if (int_out && ((int64_t)v < ilo || (int64_t)v > ihi)) inexact = true;
                dp[x * s0] = v;
            }
        }
    } else {
        // dense: taps in column-major ascending order
        int64_t klen[B2F_MAXDIM], klo[B2F_MAXDIM];
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            klen[d] = d < s->ndim ? s->len[d] : 1;
            klo[d] = d < s->ndim ? s->lo[d] : 0;
        }
        const int64_t nt = klen[0] * klen[1] * klen[2] * klen[3];
        std::vector<ACC> k(nt);
        std::vector<int64_t> koff(nt);
        {
            int64_t t = 0;
            for (int64_t j3 = 0; j3 < klen[3]; ++j3)
            for (int64_t j2 = 0; j2 < klen[2]; ++j2)
            for (int64_t j1 = 0; j1 < klen[1]; ++j1)
            for (int64_t j0 = 0; j0 < klen[0]; ++j0, ++t) {
                k[t] = (ACC)s->taps[t];
                koff[t] = (klo[0] + j0) * src.st[0] + (klo[1] + j1) * src.st[1] +
                          (klo[2] + j2) * src.st[2] + (klo[3] + j3) * src.st[3];
            }
        }
        for (int64_t i3 = R.lo[3]; i3 <= R.hi[3]; ++i3)
        for (int64_t i2 = R.lo[2]; i2 <= R.hi[2]; ++i2)
        for (int64_t i1 = R.lo[1]; i1 <= R.hi[1]; ++i1)
        for (int64_t i0 = R.lo[0]; i0 <= R.hi[0]; ++i0) {
            const S *q = src.p + src.off(i0, i1, i2, i3);
            ACC tmp = (ACC)0;
            for (int64_t t = 0; t < nt; ++t) tmp += (ACC)q[koff[t]] * k[t];
            S v;
            if (!StoreCheck<S>::conv(tmp, v)) inexact = true;
            else if (int_out && ((int64_t)v < ilo || (int64_t)v > ihi)) inexact = true;
            dst.p[dst.off(i0, i1, i2, i3)] = v;
        }
    }
    return inexact ? B2F_EINEXACT : 0;
}

template <typename S>
int dispatch_stage(const StageInfo &si, const View<S> &src, View<S> &dst, const Box &R,
                   bool int_out, int64_t ilo, int64_t ihi) {
    const int kt = si.s->tap_dtype;
    if (si.s->kind == B2F_STAGE_LAPLACIAN) return run_stage<S, S>(si, src, dst, R, int_out, ilo, ihi);
    if (std::is_same<S, double>::value) return run_stage<S, double>(si, src, dst, R, false, 0, 0);
    if (std::is_same<S, float>::value) {
        // Float32*Float32 and Float32*Int stay Float32; Float32*Float64 promotes (src/imfilter.jl:630-632)
        if (kt == B2F_TAPS_F64) return run_stage<S, double>(si, src, dst, R, false, 0, 0);
        return run_stage<S, float>(si, src, dst, R, false, 0, 0);
    }
    // integer S
    if (kt == B2F_TAPS_INT) return run_stage<S, int64_t>(si, src, dst, R, int_out, ilo, ihi);
    return run_stage<S, double>(si, src, dst, R, int_out, ilo, ihi);
}

template <typename S>
int store_out(const b2f_array *out, const View<S> &v, const Box &R) {
    // out[R] = v[R], converting S -> eltype(out) (identity except for the int64-backed integer types)
    Box oa = axes_of(out);
    View<char> ov;
    ov.init((char *)out->ptr, oa);
    const size_t es = dtype_size(out->dtype);
    for (int64_t i3 = R.lo[3]; i3 <= R.hi[3]; ++i3)
    for (int64_t i2 = R.lo[2]; i2 <= R.hi[2]; ++i2)
    for (int64_t i1 = R.lo[1]; i1 <= R.hi[1]; ++i1)
    for (int64_t i0 = R.lo[0]; i0 <= R.hi[0]; ++i0) {
        S x = v.p[v.off(i0, i1, i2, i3)];
        char *q = (char *)out->ptr + ov.off(i0, i1, i2, i3) * es;
        switch (out->dtype) {
            case B2F_F64: *(double *)q = (double)x; break;
            case B2F_F32: *(float *)q = (float)x; break;
            case B2F_U8: *(uint8_t *)q = (uint8_t)x; break;
            case B2F_I16: *(int16_t *)q = (int16_t)x; break;
            case B2F_U16: *(uint16_t *)q = (uint16_t)x; break;
            case B2F_I32: *(int32_t *)q = (int32_t)x; break;
            case B2F_U32: *(uint32_t *)q = (uint32_t)x; break;
            case B2F_I64: *(int64_t *)q = (int64_t)x; break;
            default: return fail(B2F_EARG, "unsupported output dtype %d", out->dtype);
        }
    }
    return 0;
}

struct Plan {
    int ndim;
    std::vector<StageInfo> stages;       // all stages, in order
    std::vector<int> active;             // indices of the non-copy stages
    int64_t need_lo[B2F_MAXDIM], need_hi[B2F_MAXDIM];  // padding implied by the kernel
    int64_t sum_first[B2F_MAXDIM], sum_last[B2F_MAXDIM];
    int64_t pad_lo[B2F_MAXDIM], pad_hi[B2F_MAXDIM];    // padding actually applied
    Box img_ax, out_ax, roi, padded_ax;
};

int make_plan(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int nstages,
              const b2f_border *border, const int64_t *roi_lo, const int64_t *roi_hi, Plan &P) {
    if (!img || !out || !border) return fail(B2F_EARG, "NULL argument");
    if (img->ndim < 1 || img->ndim > B2F_MAXDIM)
        return fail(B2F_ENOTSUP, "ndim %d not supported (1..%d)", img->ndim, B2F_MAXDIM);
    if (out->ndim != img->ndim)
        return fail(B2F_EDIM, "out has %d dims, img has %d", out->ndim, img->ndim);
    if (nstages < 0 || nstages > B2F_MAXSTAGES) return fail(B2F_EARG, "bad stage count %d", nstages);
    P.ndim = img->ndim;
    P.img_ax = axes_of(img);
    P.out_ax = axes_of(out);
    P.stages.resize(nstages);
    for (int d = 0; d < B2F_MAXDIM; ++d) P.sum_first[d] = P.sum_last[d] = 0;
    for (int s = 0; s < nstages; ++s) {
        int rc = analyse_stage(&stages[s], P.ndim, P.stages[s]);
        if (rc) return rc;
        if (!P.stages[s].copy) P.active.push_back(s);
        // accumulate_padding expands by every factor's axes, copy kernels included (they are 0:0)
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            P.sum_first[d] += P.stages[s].lo[d];
            P.sum_last[d] += P.stages[s].hi[d];
        }
    }
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        P.need_lo[d] = std::max<int64_t>(0, -P.sum_first[d]);  // lo(r) src/border.jl:602-606
        P.need_hi[d] = std::max<int64_t>(0, P.sum_last[d]);
        P.pad_lo[d] = P.pad_hi[d] = 0;
    }
    const int st = border->style;
    if (st < B2F_REPLICATE || st > B2F_NOPAD) return fail(B2F_EARG, "border style %d unrecognized", st);
    if (st <= B2F_FILL) {
        if (border->npad == 0) {
            for (int d = 0; d < P.ndim; ++d) { P.pad_lo[d] = P.need_lo[d]; P.pad_hi[d] = P.need_hi[d]; }
        } else if (border->npad == P.ndim) {
            for (int d = 0; d < P.ndim; ++d) {
                if (border->lo[d] < 0 || border->hi[d] < 0) return fail(B2F_EARG, "negative padding");
                P.pad_lo[d] = border->lo[d];
                P.pad_hi[d] = border->hi[d];
            }
        } else {
            return fail(B2F_EARG, "border lacks the proper padding sizes for an array with %d dimensions", P.ndim);
        }
    }
    P.padded_ax = P.img_ax;
    for (int d = 0; d < P.ndim; ++d) {
        P.padded_ax.lo[d] -= P.pad_lo[d];
        P.padded_ax.hi[d] += P.pad_hi[d];
    }
    P.roi = P.out_ax;
    if (roi_lo && roi_hi) {
        for (int d = 0; d < P.ndim; ++d) { P.roi.lo[d] = roi_lo[d]; P.roi.hi[d] = roi_hi[d]; }
    }
    return 0;
}

template <typename S>
int imfilter_typed(const b2f_array *img, const b2f_array *out, const Plan &P, const b2f_border *border) {
    const int N = P.ndim;
    int64_t ilo = 0, ihi = 0;
    const bool int_out = int_range(out->dtype, ilo, ihi);

    // (isempty(A) || isempty(kern)) && return out
    if (P.img_ax.empty() || P.roi.empty()) return 0;

    // ---- padarray(S, img, border): materialised gather copy --------------------------------
    View<S> A;
    std::vector<S> Abuf;
    {
        Box pa = P.padded_ax;
        int64_t n = 1;
        for (int d = 0; d < B2F_MAXDIM; ++d) n *= pa.len(d);
        Abuf.resize((size_t)n);
        A.init(Abuf.data(), pa);
        // per-axis source index vectors (padindices)
        std::vector<int64_t> idx[B2F_MAXDIM];
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            idx[d].resize((size_t)pa.len(d));
            const int64_t n_d = P.img_ax.len(d);
            for (int64_t t = 0; t < pa.len(d); ++t) {
                int64_t pos = t - (d < N ? P.pad_lo[d] : 0);  // 0-based position relative to img
                if (pos >= 0 && pos < n_d) { idx[d][t] = pos; continue; }
                if (border->style == B2F_FILL) { idx[d][t] = -1; continue; }
                int64_t src;
                if (!pad_source_index(border->style, pos, n_d, src))
                    return fail(B2F_EARG, "reflect padding of a length-1 axis (DivideError in the reference)");
                idx[d][t] = src;
            }
        }
        S fillv = (S)0;
        if (border->style == B2F_FILL) {
            // convert(eltype(img), value) then to S (src/borderarray.jl:11-20, src/border.jl:343)
            double fv = border->fill;
            if (img->dtype == B2F_N0F8) {
                double q = std::nearbyint(fv * 255.0);
                if (q < 0 || q > 255) return fail(B2F_EARG, "fill value %g not representable as N0f8", fv);
                uint8_t raw = (uint8_t)q;
                int rc = load_as<S>(&raw, B2F_N0F8, 0, fillv);
                if (rc) return fail(rc, "cannot convert fill");
            } else if (is_int_dtype(img->dtype)) {
                int64_t lo_, hi_;
                int_range(img->dtype, lo_, hi_);
                if (std::floor(fv) != fv || fv < (double)lo_ || fv > (double)hi_)
                    return fail(B2F_EARG, "fill value %g not representable in the image eltype", fv);
                fillv = (S)fv;
            } else if (img->dtype == B2F_F32) {
                fillv = (S)(float)fv;
            } else {
                if (std::is_integral<S>::value && std::floor(fv) != fv)
                    return fail(B2F_EINEXACT, "fill value %g not representable in eltype(out)", fv);
                fillv = (S)fv;
            }
        }
        View<char> iv;
        iv.init((char *)img->ptr, P.img_ax);
        int64_t o = 0;
        for (int64_t t3 = 0; t3 < pa.len(3); ++t3)
        for (int64_t t2 = 0; t2 < pa.len(2); ++t2)
        for (int64_t t1 = 0; t1 < pa.len(1); ++t1)
        for (int64_t t0 = 0; t0 < pa.len(0); ++t0, ++o) {
            int64_t s0 = idx[0][t0], s1 = idx[1][t1], s2 = idx[2][t2], s3 = idx[3][t3];
            if (s0 < 0 || s1 < 0 || s2 < 0 || s3 < 0) { Abuf[o] = fillv; continue; }
            int64_t lin = s0 * iv.st[0] + s1 * iv.st[1] + s2 * iv.st[2] + s3 * iv.st[3];
            S v;
            int rc = load_as<S>(img->ptr, img->dtype, lin, v);
            if (rc) return fail(rc, "cannot convert image element to eltype(out)");
            if (int_out && ((int64_t)v < ilo || (int64_t)v > ihi))
                return fail(B2F_EINEXACT, "image element does not fit eltype(out)");
            Abuf[o] = v;
        }
    }

    // ---- scheduler ---------------------------------------------------------------------------
    // inds must be inbounds for out (src/imfilter.jl:604-609)
    for (int d = 0; d < N; ++d)
        if (P.roi.lo[d] < P.out_ax.lo[d] || P.roi.hi[d] > P.out_ax.hi[d])
            return fail(B2F_EDIM, "output indices disagree with requested indices (axis %d)", d);

    const int na = (int)P.active.size();
    if (na == 0) {  // trivial kernel: copyto!(out, R, A, R)
        for (int d = 0; d < N; ++d)
            if (P.roi.lo[d] < A.ax.lo[d] || P.roi.hi[d] > A.ax.hi[d])
                return fail(B2F_EDIM, "requested indices exceed the input (axis %d)", d);
        return store_out<S>(out, A, P.roi);
    }

    // regions: R_s = roi expanded by the extents of the stages after s.  For the first stage the
    // reference computes shrink(expand(inds, calculate_padding(kernel)), k1), which is the same
    // thing; copy kernels contribute 0:0.
    std::vector<Box> R(na);
    {
        Box cur = P.roi;
        for (int a = na - 1; a >= 0; --a) {
            R[a] = cur;
            const StageInfo &si = P.stages[P.active[a]];
            for (int d = 0; d < B2F_MAXDIM; ++d) { cur.lo[d] += si.lo[d]; cur.hi[d] += si.hi[d]; }
        }
    }
    std::vector<S> T1, T2;  // ping-pong temporaries of the padded size (tempbuffer)
    View<S> src = A, dst;
    int rc_final = 0;
    for (int a = 0; a < na; ++a) {
        const StageInfo &si = P.stages[P.active[a]];
        const bool last = (a == na - 1);
        Box reg = R[a];
        // input must be big enough (src/imfilter.jl:610-615)
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            if (src.ax.lo[d] > reg.lo[d] + si.lo[d] || src.ax.hi[d] < reg.hi[d] + si.hi[d])
                return fail(B2F_EDIM,
                            "requested indices and kernel indices do not agree with indices of padded input (axis %d)", d);
        }
        std::vector<S> *tb = (a % 2 == 0) ? &T1 : &T2;
        if (tb->empty()) tb->resize(Abuf.size());
        dst.init(tb->data(), A.ax);
        for (int d = 0; d < B2F_MAXDIM; ++d)
            if (!last && (reg.lo[d] < dst.ax.lo[d] || reg.hi[d] > dst.ax.hi[d]))
                return fail(B2F_EDIM, "stage region exceeds the temporary buffer (axis %d)", d);
        if (last) {
            // write through a temporary view restricted to the region, then convert into out
            Box ob = reg;
            std::vector<S> obuf;
            int64_t n = 1;
            for (int d = 0; d < B2F_MAXDIM; ++d) n *= ob.len(d);
            obuf.resize((size_t)n);
            View<S> ov;
            ov.init(obuf.data(), ob);
            int rc = dispatch_stage<S>(si, src, ov, reg, int_out, ilo, ihi);
            if (rc == B2F_EINEXACT) rc_final = rc; else if (rc) return rc;
            rc = store_out<S>(out, ov, reg);
            if (rc) return rc;
        } else {
            int rc = dispatch_stage<S>(si, src, dst, reg, int_out, ilo, ihi);
            if (rc == B2F_EINEXACT) rc_final = rc; else if (rc) return rc;
            src = dst;
        }
    }
    if (rc_final == B2F_EINEXACT) return fail(B2F_EINEXACT, "result not representable in eltype(out) (InexactError)");
    return 0;
}

int imfilter_entry(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int nstages,
                   const b2f_border *border, const int64_t *roi_lo, const int64_t *roi_hi) {
    Plan P;
    int rc = make_plan(img, out, stages, nstages, border, roi_lo, roi_hi, P);
    if (rc) return rc;
    switch (out->dtype) {
        case B2F_F64: return imfilter_typed<double>(img, out, P, border);
        case B2F_F32: return imfilter_typed<float>(img, out, P, border);
        case B2F_U8: case B2F_I16: case B2F_U16: case B2F_I32: case B2F_U32: case B2F_I64:
            if (img->dtype == B2F_N0F8) return fail(B2F_EARG, "N0f8 image with integer output");
            return imfilter_typed<int64_t>(img, out, P, border);
    }
    return fail(B2F_EARG, "unsupported output dtype %d", out->dtype);
}

// ---- running extrema ------------------------------------------------------------------------
// Ground truth form of src/mapwindow.jl:388-481 (Lemire wedge, window [i-(w>>1), i-(w>>1)+w-1]
// truncated at the array ends) and of the generic path :270-333 for f in {minimum, maximum}:
// with any Pad style the edge buffer holds only elements of the in-image part of the window
// (copy_win! pads "win ∩ axes(img)" by padindex, :310-317), so min/max equal the truncated-window
// min/max; Fill(v) adds v wherever the window leaves the array (:326-333).
// Done separably, one axis at a time, which is exact for min/max.
template <typename T>
void extrema_axis(const T *in_mn, const T *in_mx, T *o_mn, T *o_mx, const int64_t *dims, int ax,
                  int64_t wlo, int64_t whi, bool fill, T fv) {
    int64_t pre = 1, post = 1;
    for (int d = 0; d < ax; ++d) pre *= dims[d];
    for (int d = ax + 1; d < B2F_MAXDIM; ++d) post *= dims[d];
    const int64_t n = dims[ax];
    for (int64_t q = 0; q < post; ++q)
    for (int64_t i = 0; i < n; ++i)
    for (int64_t p = 0; p < pre; ++p) {
        bool have = false, outside = false;
        T mn = T(), mx = T();
        for (int64_t j = i + wlo; j <= i + whi; ++j) {
            if (j < 0 || j >= n) { outside = true; continue; }
            const int64_t o = p + pre * (j + n * q);
            if (!have) { mn = in_mn[o]; mx = in_mx[o]; have = true; }
            else { if (in_mn[o] < mn) mn = in_mn[o]; if (in_mx[o] > mx) mx = in_mx[o]; }
        }
        if (fill && outside) {
            if (!have) { mn = mx = fv; have = true; }
            else { if (fv < mn) mn = fv; if (fv > mx) mx = fv; }
        }
        const int64_t o = p + pre * (i + n * q);
        o_mn[o] = mn;
        o_mx[o] = mx;
    }
}

template <typename T>
int extrema_typed(const b2f_array *img, const b2f_array *omin, const b2f_array *omax, int interleaved,
                  const int64_t *wlo, const int64_t *whi, const b2f_border *border) {
    const int N = img->ndim;
    int64_t dims[B2F_MAXDIM];
    int64_t n = 1;
    for (int d = 0; d < B2F_MAXDIM; ++d) { dims[d] = d < N ? img->dims[d] : 1; n *= dims[d]; }
    if (n == 0) return 0;
    const bool fill = border->style == B2F_FILL;
    T fv = (T)border->fill;
    if (fill && (double)fv != border->fill && !std::is_floating_point<T>::value)
        return fail(B2F_EARG, "fill value not representable in eltype(img)");
    if (fill && std::is_same<T, float>::value) fv = (T)(float)border->fill;
    std::vector<T> a_mn((const T *)img->ptr, (const T *)img->ptr + n), a_mx(a_mn), b_mn(n), b_mx(n);
    for (int d = 0; d < N; ++d) {
        if (wlo[d] > whi[d]) return fail(B2F_EARG, "empty window");
        if (wlo[d] == 0 && whi[d] == 0) continue;
        extrema_axis<T>(a_mn.data(), a_mx.data(), b_mn.data(), b_mx.data(), dims, d, wlo[d], whi[d], fill, fv);
        a_mn.swap(b_mn);
        a_mx.swap(b_mx);
    }
    // write the requested region (axes(out) within axes(img))
    const b2f_array *ref = omin ? omin : omax;
    Box ia = axes_of(img), oa = axes_of(ref);
    for (int d = 0; d < N; ++d) {
        if (oa.lo[d] < ia.lo[d] || oa.hi[d] > ia.hi[d]) return fail(B2F_EDIM, "output axes exceed image axes");
        if (border->style == B2F_INNER || border->style == B2F_NOPAD)
            if (oa.lo[d] + wlo[d] < ia.lo[d] || oa.hi[d] + whi[d] > ia.hi[d])
                return fail(B2F_EDIM, "output axes are not in the interior for Inner()");
    }
    View<char> iv, ov;
    iv.init(nullptr, ia);
    ov.init(nullptr, oa);
    for (int64_t i3 = oa.lo[3]; i3 <= oa.hi[3]; ++i3)
    for (int64_t i2 = oa.lo[2]; i2 <= oa.hi[2]; ++i2)
    for (int64_t i1 = oa.lo[1]; i1 <= oa.hi[1]; ++i1)
    for (int64_t i0 = oa.lo[0]; i0 <= oa.hi[0]; ++i0) {
        const int64_t si = iv.off(i0, i1, i2, i3), oi = ov.off(i0, i1, i2, i3);
        if (interleaved) {
            ((T *)omin->ptr)[2 * oi] = a_mn[si];
            ((T *)omin->ptr)[2 * oi + 1] = a_mx[si];
        } else {
            if (omin) ((T *)omin->ptr)[oi] = a_mn[si];
            if (omax) ((T *)omax->ptr)[oi] = a_mx[si];
        }
    }
    return 0;
}

}  // namespace

extern "C" {

const char *b2f_version(void) { return "b2f-oracle 0.1 (CPU restatement of ImageFiltering.jl v0.7.12 FIR path)"; }
const char *b2f_last_error(void) { return g_err.c_str(); }
int b2f_is_device_library(void) { return 0; }

int b2f_set_device(int) { return 0; }
// the oracle IS the exact arithmetic: the mode is accepted and ignored (B2F_ACCUM_FMA is a permission)
int b2f_set_accum_mode(int32_t mode) { return (mode == 0 || mode == 1) ? 0 : fail(B2F_EARG, "unknown accumulate mode"); }
int b2f_device_count(int *count) { if (count) *count = 0; return 0; }
int b2f_sm_count(int *count) { if (count) *count = 0; return 0; }
int b2f_bench_fma_peak(double *tfma_per_s, void *) { if (tfma_per_s) *tfma_per_s = 0.0; return 0; }
int b2f_malloc(void **, uint64_t) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_free(void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_host_alloc(void **, uint64_t) { return fail(B2F_ENOTSUP, "oracle library has no pinned memory"); }
int b2f_host_free(void *) { return fail(B2F_ENOTSUP, "oracle library has no pinned memory"); }
int b2f_host_register(void *, uint64_t) { return fail(B2F_ENOTSUP, "oracle library has no pinned memory"); }
int b2f_host_unregister(void *) { return fail(B2F_ENOTSUP, "oracle library has no pinned memory"); }
int b2f_memcpy_h2d(void *, const void *, uint64_t) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_memcpy_d2h(void *, const void *, uint64_t) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_sync(void) { return 0; }
int b2f_ipc_export(const void *, void *, uint64_t *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_ipc_open(const void *, uint64_t, void **) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_ipc_close(void *, uint64_t) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int64_t b2f_launch_count(void) { return 0; }
void b2f_reset_launch_count(void) {}
const char *b2f_last_path(void) { return "oracle"; }

int b2f_imfilter(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                 const b2f_border *border, const int64_t *roi_lo, const int64_t *roi_hi, void *) {
    return imfilter_entry(img, out, stages, nstages, border, roi_lo, roi_hi);
}

int b2f_imgradients(const b2f_array *img, const b2f_array *outs, int32_t nplanes,
                    const b2f_stage *stages, int32_t nstages_each, const b2f_border *border, void *) {
    // src/specialty.jl:47-51: one independent imfilter call per plane
    for (int p = 0; p < nplanes; ++p) {
        int rc = imfilter_entry(img, &outs[p], stages + (size_t)p * nstages_each, nstages_each, border,
                                nullptr, nullptr);
        if (rc) return rc;
    }
    return 0;
}

int b2f_mapwindow_extrema(const b2f_array *img, const b2f_array *out_min, const b2f_array *out_max,
                          int32_t interleaved, const int64_t *win_lo, const int64_t *win_hi,
                          const b2f_border *border, void *) {
    if (!img || !border || !win_lo || !win_hi || (!out_min && !out_max))
        return fail(B2F_EARG, "NULL argument");
    if (interleaved && !out_min) return fail(B2F_EARG, "interleaved output needs out_min");
    if (img->ndim < 1 || img->ndim > B2F_MAXDIM) return fail(B2F_ENOTSUP, "ndim %d not supported", img->ndim);
    if (border->style == B2F_INNER)
        for (int d = 0; d < img->ndim; ++d)
            if (win_hi[d] - win_lo[d] + 1 > img->dims[d])
                return fail(B2F_EDIM, "window is larger than the image along axis %d: no interior for Inner()", d);
    const b2f_array *ref = out_min ? out_min : out_max;
    if (ref->dtype != img->dtype || (out_max && out_max->dtype != img->dtype))
        return fail(B2F_EARG, "extrema outputs must have the image eltype");
    switch (img->dtype) {
        case B2F_F32: return extrema_typed<float>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
        case B2F_F64: return extrema_typed<double>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
        case B2F_U8: case B2F_N0F8:
            return extrema_typed<uint8_t>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
        case B2F_I16: return extrema_typed<int16_t>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
        case B2F_U16: return extrema_typed<uint16_t>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
        case B2F_I32: return extrema_typed<int32_t>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
        case B2F_U32: return extrema_typed<uint32_t>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
        case B2F_I64: return extrema_typed<int64_t>(img, out_min, out_max, interleaved, win_lo, win_hi, border);
    }
    return fail(B2F_EARG, "unsupported dtype");
}

int b2f_oracle_padarray(const b2f_array *img, const b2f_array *out, const b2f_border *border);

// Slab form (include/b2f.h): the owned planes of imfilter on the whole array.  Restated as: materialise exactly the
// planes the owned outputs read along the sharded axis — each taken from the owned slab, from a halo buffer (matched
// on its logical index), through the global border remap (padindex, src/border.jl:564-596), or the Fill value —
// then run the ordinary cascade with `out` = the owned planes; along the sharded axis every read then stays inside
// the materialised planes, so only the other axes see the border.
int b2f_imfilter_slab(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                      const b2f_border *border, int64_t global_last_dim, int64_t slab_first,
                      const void *halo_lo, int64_t n_halo_lo, const void *halo_hi, int64_t n_halo_hi, void *) {
    if (!img || !out || !border || !stages) return fail(B2F_EARG, "NULL argument");
    const int N = img->ndim, last = N - 1;
    if (N < 2 || N > B2F_MAXDIM || out->ndim != N) return fail(B2F_EDIM, "slab arrays need 2..4 dims and equal rank");
    if (border->style > B2F_FILL) return fail(B2F_ENOTSUP, "slab form supports Pad and Fill borders");
    const int64_t own_n = img->dims[last];
    if (own_n < 1) return 0;
    if (slab_first < 0 || slab_first + own_n > global_last_dim) return fail(B2F_EDIM, "slab lies outside the global axis");
    int64_t zlo = 0, zhi = 0;
    for (int s = 0; s < nstages; ++s) {
        if (stages[s].kind != B2F_STAGE_1D) return fail(B2F_ENOTSUP, "slab form takes 1-D stages");
        if (stages[s].axis == last) { zlo += stages[s].lo[last]; zhi += stages[s].lo[last] + stages[s].len[last] - 1; }
    }
    if (zlo > 0) zlo = 0;
    if (zhi < 0) zhi = 0;
    int64_t plane = 1;
    for (int d = 0; d < last; ++d) plane *= img->dims[d];
    const size_t es = dtype_size(img->dtype), pbytes = (size_t)plane * es;
    const int64_t nz = own_n - zlo + zhi;
    std::vector<unsigned char> ext((size_t)nz * pbytes);
    // Fill planes: the value converted through eltype(img), like padarray does (src/borderarray.jl:11-20)
    std::vector<unsigned char> fillplane;
    for (int64_t k = 0; k < nz; ++k) {
        const int64_t z = slab_first + zlo + k;   // logical plane index
        const unsigned char *src = nullptr;
        auto locate = [&](int64_t g) -> const unsigned char * {
            if (g >= slab_first && g < slab_first + own_n) return (const unsigned char *)img->ptr + (size_t)(g - slab_first) * pbytes;
            if (g < slab_first && g >= slab_first - n_halo_lo) return (const unsigned char *)halo_lo + (size_t)(g - (slab_first - n_halo_lo)) * pbytes;
            if (g >= slab_first + own_n && g < slab_first + own_n + n_halo_hi) return (const unsigned char *)halo_hi + (size_t)(g - slab_first - own_n) * pbytes;
            return nullptr;
        };
        src = locate(z);
        bool is_fill = false;
        if (!src) {
            if (z >= 0 && z < global_last_dim) return fail(B2F_EDIM, "halo too small: plane %lld is needed but not present", (long long)z);
            if (border->style == B2F_FILL) {
                is_fill = true;
            } else {
                int64_t g = 0;
                if (!pad_source_index(border->style, z, global_last_dim, g)) return fail(B2F_EARG, "reflect padding of a length-1 axis");
                src = locate(g);
                if (!src) return fail(B2F_EDIM, "halo too small: plane %lld is needed but not present", (long long)g);
            }
        }
        unsigned char *dst = ext.data() + (size_t)k * pbytes;
        if (!is_fill) { memcpy(dst, src, pbytes); continue; }
        if (fillplane.empty()) {   // one padded plane of a 1-plane dummy gives the converted fill value
            fillplane.resize(pbytes);
            b2f_array one = *img, pad1 = *img;
            one.ptr = (void *)img->ptr; one.dims[last] = 1;
            std::vector<unsigned char> two(2 * pbytes);
            pad1.ptr = two.data(); pad1.dims[last] = 2;
            b2f_border fb = *border;
            fb.npad = N;
            for (int d = 0; d < B2F_MAXDIM; ++d) fb.lo[d] = fb.hi[d] = 0;
            fb.hi[last] = 1;
            int rc = b2f_oracle_padarray(&one, &pad1, &fb);
            if (rc) return rc;
            memcpy(fillplane.data(), two.data() + pbytes, pbytes);
        }
        memcpy(dst, fillplane.data(), pbytes);
    }
    b2f_array e = *img, o = *out;
    e.ptr = ext.data();
    e.dims[last] = nz;
    e.origin[last] = img->origin[last] + zlo;     // the slab's own first plane keeps index img->origin[last]
    o.origin[last] = img->origin[last];
    for (int d = 0; d < last; ++d) o.origin[d] = img->origin[d];
    e.mem = o.mem = B2F_HOST;
    return imfilter_entry(&e, &o, stages, nstages, border, nullptr, nullptr);
}

// ---- oracle-only entry points (not in b2f.h) ---------------------------------------------------

// padarray(S, img, border) with explicit lo/hi — exposes the border restatement to the golden tests
// (reference test/border.jl:35-206).  out must have dims img.dims + lo + hi; eltype(out) = S.
int b2f_oracle_padarray(const b2f_array *img, const b2f_array *out, const b2f_border *border) {
    // a cascade of zero stages == copy of the padded array over axes(out)
    if (!img || !out || !border) return fail(B2F_EARG, "NULL argument");
    return imfilter_entry(img, out, nullptr, 0, border, nullptr, nullptr);
}

// The Lemire streaming max-min of src/mapwindow.jl:426-473 restated with explicit deques, along
// axis 0 of a (n, m) column-major array of (min,max) pairs; used by the tests to pin the separable
// ground-truth form above against the reference's actual algorithm (window placement, strict
// comparisons, delayed write-back).
int b2f_oracle_lemire_axis0(double *mn, double *mx, int64_t n, int64_t m, int64_t window) {
    if (window < 1) return fail(B2F_EARG, "window must be positive");
    if (window == 1 || n == 0) return 0;
    const int64_t half = window >> 1;
    std::vector<int64_t> L(window + 2), U(window + 2);
    std::vector<double> cmn(std::max<int64_t>(half, 1)), cmx(std::max<int64_t>(half, 1));
    for (int64_t t = 0; t < half; ++t) { cmn[t] = mn[0]; cmx[t] = mx[0]; }  // cache = ntuple(i -> first(A), w>>1)
    int64_t ch = 0;                    // the cache and `c` persist across columns, as in the reference
    double c0 = mn[0], c1 = mx[0];
    for (int64_t J = 0; J < m; ++J) {
        double *a = mn + J * n, *b = mx + J * n;
        int64_t Lh = 0, Lt = 0, Uh = 0, Ut = 0;  // deques as [head, tail) over a ring of window+2
        const int64_t cap = window + 2;
        auto Lfirst = [&]() { return L[Lh % cap]; };
        auto Ufirst = [&]() { return U[Uh % cap]; };
        auto addtoback = [&](int64_t i) {
            while (Lt > Lh && a[i] < a[L[(Lt - 1) % cap]]) --Lt;
            while (Ut > Uh && b[i] > b[U[(Ut - 1) % cap]]) --Ut;
            L[Lt % cap] = i; ++Lt;
            U[Ut % cap] = i; ++Ut;
        };
        auto cycle = [&](double x0, double x1) {
            if (half == 0) { c0 = x0; c1 = x1; return; }
            c0 = cmn[ch]; c1 = cmx[ch];
            cmn[ch] = x0; cmx[ch] = x1;
            ch = (ch + 1) % half;
        };
        const int64_t iw = std::min(n - 1, window - 1);
        for (int64_t i = 0; i <= iw; ++i) { addtoback(i); cycle(a[Lfirst()], b[Ufirst()]); }
        for (int64_t i = iw + 1; i <= n - 1; ++i) {
            // A[i-window] = c is safe to overwrite: i-window < every index still in the wedges
            a[i - window] = c0; b[i - window] = c1;
            if (i == window + Ufirst()) ++Uh;
            if (i == window + Lfirst()) ++Lh;
            addtoback(i);
            cycle(a[Lfirst()], b[Ufirst()]);
        }
        for (int64_t i = n - window; i <= n - 2; ++i) {
            if (i >= 0) { a[i] = c0; b[i] = c1; }
            if (i == Ufirst()) ++Uh;
            if (i == Lfirst()) ++Lh;
            cycle(a[Lfirst()], b[Ufirst()]);
        }
        a[n - 1] = c0; b[n - 1] = c1;
    }
    return 0;
}

// CPUThreads(Algorithm.FIRTiled(tilesize)) for cascades of >= 2 stages (src/imfilter.jl:398-404,
// 476-542): serial pad copy, tile list statically partitioned over `nthreads` workers, one scratch
// tile pair per worker whose eltype is filter_type (== eltype(out) here).  Per-pixel arithmetic is
// the same as the untiled path.  Supports float outputs and 1-D stages (the CPU-baseline use).
// `tile` = TiledIteration.padded_tilesize result; that package is not under /root/reference, so the
// caller passes the tile size (bench.py uses tiles of about 32 KiB, the L1-sized choice the
// package documents).
int b2f_oracle_imfilter_tiled(const b2f_array *img, const b2f_array *out, const b2f_stage *stages,
                              int32_t nstages, const b2f_border *border, const int64_t *tile,
                              int32_t nthreads);

}  // extern "C"

namespace {

template <typename S>
int tiled_typed(const b2f_array *img, const b2f_array *out, const Plan &P, const b2f_border *border,
                const int64_t *tile, int nthreads) {
    const int N = P.ndim;
    const int na = (int)P.active.size();
    if (na < 2) return fail(B2F_ENOTSUP, "tiled path needs a cascade of >= 2 stages");
    // pad (serial, as in the reference)
    b2f_array padded = *img;
    Box pa = P.padded_ax;
    int64_t n = 1;
    for (int d = 0; d < B2F_MAXDIM; ++d) n *= pa.len(d);
    std::vector<S> Abuf((size_t)n);
    {
        b2f_array tmp;
        memset(&tmp, 0, sizeof tmp);
        tmp.ptr = Abuf.data();
        tmp.dtype = std::is_same<S, float>::value ? B2F_F32 : B2F_F64;
        tmp.ndim = N;
        for (int d = 0; d < N; ++d) { tmp.dims[d] = pa.len(d); tmp.origin[d] = pa.lo[d]; }
        b2f_border b = *border;
        b.npad = N;
        for (int d = 0; d < N; ++d) { b.lo[d] = P.pad_lo[d]; b.hi[d] = P.pad_hi[d]; }
        int rc = imfilter_entry(img, &tmp, nullptr, 0, &b, nullptr, nullptr);
        if (rc) return rc;
    }
    (void)padded;
    View<S> A;
    A.init(Abuf.data(), pa);
    // extents of the stages after the first (kt)
    int64_t kt_lo[B2F_MAXDIM] = {0, 0, 0, 0}, kt_hi[B2F_MAXDIM] = {0, 0, 0, 0};
    for (int a = 1; a < na; ++a)
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            kt_lo[d] += P.stages[P.active[a]].lo[d];
            kt_hi[d] += P.stages[P.active[a]].hi[d];
        }
    int64_t chunk[B2F_MAXDIM];
    int64_t tsz[B2F_MAXDIM];
    int64_t tcount = 1;
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        tsz[d] = d < N ? tile[d] : 1;
        chunk[d] = tsz[d] - (kt_hi[d] - kt_lo[d]);
        if (chunk[d] < 1) return fail(B2F_EARG, "tile smaller than the kernel");
        tcount *= tsz[d];
    }
    // tile list over roi
    std::vector<Box> tiles;
    {
        Box r = P.roi;
        for (int64_t c3 = r.lo[3]; c3 <= r.hi[3]; c3 += chunk[3])
        for (int64_t c2 = r.lo[2]; c2 <= r.hi[2]; c2 += chunk[2])
        for (int64_t c1 = r.lo[1]; c1 <= r.hi[1]; c1 += chunk[1])
        for (int64_t c0 = r.lo[0]; c0 <= r.hi[0]; c0 += chunk[0]) {
            Box t;
            int64_t c[4] = {c0, c1, c2, c3};
            for (int d = 0; d < B2F_MAXDIM; ++d) {
                t.lo[d] = c[d];
                t.hi[d] = std::min(c[d] + chunk[d] - 1, r.hi[d]);
            }
            tiles.push_back(t);
        }
    }
    const int64_t nt = (int64_t)tiles.size();
    const int T = std::max(1, nthreads);
    const int64_t per = (nt + T - 1) / T;  // Iterators.partition(tileinds_all, ceil(Int, n/ntasks))
    std::atomic<int> bad(0);
    auto worker = [&](int w) {
        std::vector<S> b1((size_t)tcount), b2((size_t)tcount);
        for (int64_t t = w * per; t < std::min(nt, (w + 1) * per); ++t) {
            Box tinds = tiles[t];
            Box cur = tinds;  // expand(tinds, kt)
            for (int d = 0; d < B2F_MAXDIM; ++d) { cur.lo[d] += kt_lo[d]; cur.hi[d] += kt_hi[d]; }
            View<S> src = A, dst;
            std::vector<S> *bufs[2] = {&b1, &b2};
            for (int a = 0; a < na; ++a) {
                const StageInfo &si = P.stages[P.active[a]];
                if (a == na - 1) {
                    std::vector<S> obuf;
                    int64_t m = 1;
                    for (int d = 0; d < B2F_MAXDIM; ++d) m *= cur.len(d);
                    obuf.resize((size_t)m);
                    View<S> ov;
                    ov.init(obuf.data(), cur);
                    if (dispatch_stage<S>(si, src, ov, cur, false, 0, 0)) bad = 1;
                    if (store_out<S>(out, ov, cur)) bad = 1;
                } else {
                    dst.init(bufs[a & 1]->data(), cur);  // TileBuffer(tile, tileinds)
                    if (dispatch_stage<S>(si, src, dst, cur, false, 0, 0)) bad = 1;
                    src = dst;
                    const StageInfo &nx = P.stages[P.active[a + 1]];
                    for (int d = 0; d < B2F_MAXDIM; ++d) { cur.lo[d] -= nx.lo[d]; cur.hi[d] -= nx.hi[d]; }
                }
            }
        }
    };
    if (T == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < T; ++w) th.emplace_back(worker, w);
        for (auto &t : th) t.join();
    }
    return bad ? fail(B2F_EARG, "tiled stage failed") : 0;
}

}  // namespace

extern "C" int b2f_oracle_imfilter_tiled(const b2f_array *img, const b2f_array *out, const b2f_stage *stages,
                                         int32_t nstages, const b2f_border *border, const int64_t *tile,
                                         int32_t nthreads) {
    Plan P;
    int rc = make_plan(img, out, stages, nstages, border, nullptr, nullptr, P);
    if (rc) return rc;
    if (border->style > B2F_FILL) return fail(B2F_ENOTSUP, "tiled oracle handles Pad/Fill borders");
    if (out->dtype == B2F_F64) return tiled_typed<double>(img, out, P, border, tile, nthreads);
    if (out->dtype == B2F_F32) return tiled_typed<float>(img, out, P, border, tile, nthreads);
    return fail(B2F_ENOTSUP, "tiled oracle handles float outputs");
}

// ---- §8(f) rank 1: findlocalextrema / blob_LoG plumbing (reference src/extrema.jl:85-105, 125-164) ----------------
namespace {
double o_elem(const b2f_array *a, int64_t i) {          // element as Float64 (N0f8: the raw code, same order)
    switch (a->dtype) {
        case B2F_U8: case B2F_N0F8: return (double)((const uint8_t *)a->ptr)[i];
        case B2F_I16: return (double)((const int16_t *)a->ptr)[i];
        case B2F_U16: return (double)((const uint16_t *)a->ptr)[i];
        case B2F_I32: return (double)((const int32_t *)a->ptr)[i];
        case B2F_U32: return (double)((const uint32_t *)a->ptr)[i];
        case B2F_I64: return (double)((const int64_t *)a->ptr)[i];
        case B2F_F32: return (double)((const float *)a->ptr)[i];
        default: return ((const double *)a->ptr)[i];
    }
}
int64_t o_numel(const b2f_array *a) {
    int64_t n = 1;
    for (int d = 0; d < a->ndim; ++d) n *= a->dims[d] < 0 ? 0 : a->dims[d];
    return n;
}
}  // namespace

extern "C" {

// src/extrema.jl:125-162: `for i in R` (column-major) over the edge-clipped indices; i is an extremum when f(img[i],
// img[i+j]) holds for every window offset j != 0 whose target lies inside the array
int b2f_findlocalextrema(const b2f_array *img, int32_t minima, const int64_t *window, const int32_t *edges,
                         int64_t *idx, int64_t cap, int64_t *count, void *) {
    if (!img || !window || !edges || !count || (cap > 0 && !idx)) return fail(B2F_EARG, "NULL argument");
    if (img->ndim < 1 || img->ndim > B2F_MAXDIM) return fail(B2F_ENOTSUP, "ndim %d not supported", img->ndim);
    int64_t dims[B2F_MAXDIM], half[B2F_MAXDIM], clip[B2F_MAXDIM], stride[B2F_MAXDIM];
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        dims[d] = d < img->ndim ? img->dims[d] : 1;
        half[d] = clip[d] = 0;
        if (d < img->ndim) {
            if (window[d] < 1) return fail(B2F_EARG, "window sizes must be positive");
            half[d] = window[d] >> 1;
            clip[d] = edges[d] ? 0 : 1;
        }
        stride[d] = d == 0 ? 1 : stride[d - 1] * dims[d - 1];
    }
    const int64_t n = o_numel(img);
    const bool is_i64 = img->dtype == B2F_I64;
    int64_t found = 0;
    for (int64_t lin = 0; lin < n; ++lin) {
        int64_t c[B2F_MAXDIM], r = lin;
        bool ok = true;
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            c[d] = r % dims[d];
            r /= dims[d];
            if (clip[d] && (c[d] == 0 || c[d] == dims[d] - 1)) ok = false;
        }
        if (!ok) continue;
        const double v = o_elem(img, lin);
        const int64_t vi = is_i64 ? ((const int64_t *)img->ptr)[lin] : 0;
        for (int64_t j3 = -half[3]; j3 <= half[3] && ok; ++j3)
            for (int64_t j2 = -half[2]; j2 <= half[2] && ok; ++j2)
                for (int64_t j1 = -half[1]; j1 <= half[1] && ok; ++j1)
                    for (int64_t j0 = -half[0]; j0 <= half[0] && ok; ++j0) {
                        if (!j0 && !j1 && !j2 && !j3) continue;
                        const int64_t q[4] = {c[0] + j0, c[1] + j1, c[2] + j2, c[3] + j3};
                        bool inside = true;
                        for (int d = 0; d < 4; ++d) inside = inside && q[d] >= 0 && q[d] < dims[d];
                        if (!inside) continue;
                        const int64_t t = lin + j0 * stride[0] + j1 * stride[1] + j2 * stride[2] + j3 * stride[3];
                        if (is_i64) {
                            const int64_t w = ((const int64_t *)img->ptr)[t];
                            ok = minima ? vi < w : vi > w;
                        } else {
                            const double w = o_elem(img, t);
                            ok = minima ? v < w : v > w;
                        }
                    }
        if (ok) {
            if (found < cap) idx[found] = lin;
            ++found;
        }
    }
    *count = found;
    return 0;
}

// src/extrema.jl:99-103: the slice of the leading axis receives the filtered image, then `.*= -σ`
int b2f_scale_into_slice(const b2f_array *src, const b2f_array *stack, int64_t slice, double scale, void *) {
    if (!src || !stack) return fail(B2F_EARG, "NULL argument");
    if (stack->ndim != src->ndim + 1 || stack->ndim > B2F_MAXDIM) return fail(B2F_EDIM, "stack must have one leading axis more than src");
    for (int d = 0; d < src->ndim; ++d)
        if (stack->dims[d + 1] != src->dims[d]) return fail(B2F_EDIM, "stack axes do not match src axes");
    if (slice < 0 || slice >= stack->dims[0]) return fail(B2F_EDIM, "slice outside the leading axis");
    if (stack->dtype != B2F_F32 && stack->dtype != B2F_F64) return fail(B2F_EARG, "stack must be Float32 or Float64");
    const int64_t n = o_numel(src), S = stack->dims[0];
    for (int64_t i = 0; i < n; ++i) {
        const double v = o_elem(src, i) * scale;
        if (stack->dtype == B2F_F32) ((float *)stack->ptr)[slice + S * i] = (float)v;
        else ((double *)stack->ptr)[slice + S * i] = v;
    }
    return 0;
}

// src/extrema.jl:85: imgmax = maximum(abs, img)
int b2f_maxabs(const b2f_array *img, double *result, void *) {
    if (!img || !result) return fail(B2F_EARG, "NULL argument");
    const int64_t n = o_numel(img);
    if (n == 0) return fail(B2F_EARG, "reducing over an empty collection is not allowed");
    double m = 0.0;
    bool nan = false;
    for (int64_t i = 0; i < n; ++i) {
        const double v = std::fabs(o_elem(img, i));
        if (v != v) nan = true; else if (v > m) m = v;
    }
    *result = nan ? std::numeric_limits<double>::quiet_NaN() : m;
    return 0;
}

// src/extrema.jl:86-90: img_LoG[x] for every peak x
int b2f_gather(const b2f_array *arr, const int64_t *idx, int64_t n, double *values, void *) {
    if (!arr || (n > 0 && (!idx || !values))) return fail(B2F_EARG, "NULL argument");
    const int64_t total = o_numel(arr);
    for (int64_t i = 0; i < n; ++i) {
        if (idx[i] < 0 || idx[i] >= total) return fail(B2F_EDIM, "index outside the array");
        values[i] = o_elem(arr, idx[i]);
    }
    return 0;
}

}  // extern "C"

// ---- NA() border pieces (reference src/imfilter.jl:282-318, 1110-1127, 1234-1250) ---------------------------------
extern "C" {

int b2f_na_prepare(const b2f_array *img, int32_t na_mode, const b2f_array *imgtmp, const b2f_array *valid, int32_t *hasna, void *) {
    if (!img || !hasna) return fail(B2F_EARG, "NULL argument");
    if (na_mode < 0 || na_mode > 2) return fail(B2F_EARG, "na_mode must be 0 (isnan), 1 (!isfinite) or 2 (never)");
    const int64_t n = o_numel(img);
    for (const b2f_array *a : {imgtmp, valid}) {
        if (!a) continue;
        if (a->dtype != B2F_F32 && a->dtype != B2F_F64) return fail(B2F_EARG, "imgtmp / valid must be Float32 or Float64");
        if (o_numel(a) != n) return fail(B2F_EDIM, "imgtmp / valid must have the axes of img");
    }
    int any = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double v = o_elem(img, i);
        const bool na = na_mode == 0 ? (v != v) : na_mode == 1 ? !std::isfinite(v) : false;
        any |= na ? 1 : 0;
        if (imgtmp) {
            if (imgtmp->dtype == B2F_F32)
                ((float *)imgtmp->ptr)[i] = na ? 0.f : (img->dtype == B2F_N0F8 ? (float)((const uint8_t *)img->ptr)[i] / 255.0f : (float)v);
            else
                ((double *)imgtmp->ptr)[i] = na ? 0.0 : (img->dtype == B2F_N0F8 ? v / 255.0 : v);
        }
        if (valid) {
            if (valid->dtype == B2F_F32) ((float *)valid->ptr)[i] = na ? 0.f : 1.f;
            else ((double *)valid->ptr)[i] = na ? 0.0 : 1.0;
        }
    }
    *hasna = any;
    return 0;
}

int b2f_divide(const b2f_array *out, const b2f_array *den, void *) {
    if (!out || !den) return fail(B2F_EARG, "NULL argument");
    if ((out->dtype != B2F_F32 && out->dtype != B2F_F64) || (den->dtype != B2F_F32 && den->dtype != B2F_F64))
        return fail(B2F_ENOTSUP, "divide takes Float32 / Float64 arrays");
    const int64_t n = o_numel(out);
    if (o_numel(den) != n) return fail(B2F_EDIM, "out and den must have the same axes");
    for (int64_t i = 0; i < n; ++i) {
        if (out->dtype == B2F_F32 && den->dtype == B2F_F32) ((float *)out->ptr)[i] = ((float *)out->ptr)[i] / ((const float *)den->ptr)[i];
        else {
            const double q = o_elem(out, i) / o_elem(den, i);
            if (out->dtype == B2F_F32) ((float *)out->ptr)[i] = (float)q; else ((double *)out->ptr)[i] = q;
        }
    }
    return 0;
}

int b2f_normalize_dims(const b2f_array *out, const double *const *factors, void *) {
    if (!out || !factors) return fail(B2F_EARG, "NULL argument");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_ENOTSUP, "normalize_dims takes a Float32 / Float64 array");
    const int64_t n = o_numel(out);
    for (int64_t i = 0; i < n; ++i) {
        int64_t r = i;
        double t = o_elem(out, i);
        for (int d = 0; d < out->ndim; ++d) {
            const int64_t c = r % out->dims[d];
            r /= out->dims[d];
            t = t / factors[d][c];
        }
        if (out->dtype == B2F_F32) ((float *)out->ptr)[i] = (float)t; else ((double *)out->ptr)[i] = t;
    }
    return 0;
}

}  // extern "C"

// staged slab form / stream-ordered copies: device-only features; the oracle exports the symbols and declines
extern "C" {
int b2f_imfilter_slab_staged(const b2f_array *, const b2f_array *, const b2f_stage *, int32_t, const b2f_border *, int64_t, int64_t,
                             const void *, int64_t, const void *, int64_t, const void *, const void *, int32_t, int32_t, void *) {
    return fail(B2F_ENOTSUP, "oracle library has no copy engines: use b2f_imfilter_slab");
}
int b2f_imfilter_slab_xy(const b2f_array *, const b2f_array *, const b2f_stage *, int32_t, const b2f_border *, int64_t, int64_t,
                         const b2f_slab_xy *, const void *, const void *, int32_t, int32_t, void *) {
    return fail(B2F_ENOTSUP, "oracle library: the xy-filtered halo form belongs to the fused device kernel; use b2f_imfilter_slab");
}
int b2f_shard_ctx_create(b2f_shard_ctx **, int32_t, int32_t) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_shard_ctx_export(b2f_shard_ctx *, const void *, int64_t, void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_shard_ctx_connect(b2f_shard_ctx *, const void *, const void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_shard_handshake(b2f_shard_ctx *, void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_imfilter_sharded(b2f_shard_ctx *, const b2f_array *, const b2f_array *, const b2f_stage *, int32_t, const b2f_border *, int64_t, int64_t,
                         void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_shard_ctx_destroy(b2f_shard_ctx *) { return 0; }
int b2f_memcpy_async(void *, const void *, uint64_t, void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_memcpy2d_async(void *, uint64_t, const void *, uint64_t, uint64_t, uint64_t, void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_stream_write32(void *, uint32_t, void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_stream_wait_geq32(void *, uint32_t, void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
int b2f_memset_async(void *, int32_t, uint64_t, void *) { return fail(B2F_ENOTSUP, "oracle library has no device memory"); }
}

// ---- mapwindow(median!, ...) (reference src/mapwindow.jl:270-333 with f = median!; Statistics.median!) ------------------
namespace {
int64_t o_remap(int style, int64_t i, int64_t n) {        // src/border.jl:564-590 relative to a range of length n
    if (i >= 0 && i < n) return i;
    switch (style) {
        case B2F_REPLICATE: return i < 0 ? 0 : n - 1;
        case B2F_CIRCULAR: { int64_t m = i % n; return m < 0 ? m + n : m; }
        case B2F_SYMMETRIC: { const int64_t p = 2 * n; int64_t m = i % p; if (m < 0) m += p; return m < n ? m : p - 1 - m; }
        case B2F_REFLECT: { const int64_t p = 2 * n - 2; int64_t m = i % p; if (m < 0) m += p; return m < n ? m : p - m; }
        default: return -1;
    }
}
int64_t o_win_index(int style, int64_t k, int64_t a, int64_t b, int64_t n) {   // copy_win!: padindex on window ∩ image
    if (style == B2F_FILL) return (k >= 0 && k < n) ? k : -1;
    const int64_t lo = a > 0 ? a : 0, hi = b < n - 1 ? b : n - 1, len = hi - lo + 1;
    if (len == 1) return lo;
    return lo + o_remap(style, k - lo, len);
}
}  // namespace

static void o_store_int(const b2f_array *a, int64_t i, int64_t v) {
    switch (a->dtype) {
        case B2F_U8: ((uint8_t *)a->ptr)[i] = (uint8_t)v; break;
        case B2F_I16: ((int16_t *)a->ptr)[i] = (int16_t)v; break;
        case B2F_U16: ((uint16_t *)a->ptr)[i] = (uint16_t)v; break;
        case B2F_I32: ((int32_t *)a->ptr)[i] = (int32_t)v; break;
        case B2F_U32: ((uint32_t *)a->ptr)[i] = (uint32_t)v; break;
        default: ((int64_t *)a->ptr)[i] = v; break;
    }
}

// The generic window loop of the reference (src/mapwindow.jl:270-306) for f in {median!, mean, sum, minimum, maximum} with
// `indices=` ranges (:123-131,156-183): output element j along axis d is the window at image index idx_first[d] + j*idx_step[d].
// mean / sum reduce the window in its memory order with the reference's accumulator (Float32 windows in Float32, integers in Int).
static int o_mapwindow_reduce(const b2f_array *img, const b2f_array *out, int op, const int64_t *win_lo, const int64_t *win_hi,
                              const b2f_border *border, const int64_t *idx_first, const int64_t *idx_step) {
    if (!img || !out || !win_lo || !win_hi || !border) return fail(B2F_EARG, "NULL argument");
    if ((idx_first == nullptr) != (idx_step == nullptr)) return fail(B2F_EARG, "idx_first and idx_step go together");
    if (op < 0 || op > 4) return fail(B2F_EARG, "unknown window reduction %d", op);
    const int N = img->ndim;
    if (N < 1 || N > B2F_MAXDIM || out->ndim != N) return fail(B2F_EDIM, "mapwindow needs 1..4 dims and equal rank");
    if (img->dtype == B2F_N0F8) return fail(B2F_ENOTSUP, "window reductions of N0f8 images are not available");
    const bool isf = img->dtype == B2F_F32 || img->dtype == B2F_F64;
    int want;
    if (op == 0 || op == 1) want = img->dtype == B2F_F32 ? B2F_F32 : B2F_F64;
    else if (op == 2) want = isf ? img->dtype : B2F_I64;
    else want = img->dtype;
    if (out->dtype != want) return fail(B2F_EARG, "output eltype %d does not match the reduction's result type %d", out->dtype, want);
    if (border->style > B2F_INNER) return fail(B2F_ENOTSUP, "border style %d is not supported by mapwindow", border->style);
    const int style = border->style == B2F_INNER ? B2F_REPLICATE : border->style;
    int64_t dims[4], odims[4], ooff[4], ostep[4], wlo[4], wn[4], stride[4], nout = 1, wtotal = 1;
    for (int d = 0; d < 4; ++d) {
        dims[d] = d < N ? img->dims[d] : 1;
        odims[d] = d < N ? out->dims[d] : 1;
        ostep[d] = (d < N && idx_step) ? idx_step[d] : 1;
        ooff[d] = d < N ? ((idx_first ? idx_first[d] : out->origin[d]) - img->origin[d]) : 0;
        wlo[d] = d < N ? win_lo[d] : 0;
        wn[d] = d < N ? win_hi[d] - win_lo[d] + 1 : 1;
        if (wn[d] < 1) return fail(B2F_EARG, "empty window");
        if (ostep[d] < 1) return fail(B2F_EARG, "indices must be increasing ranges");
        wtotal *= wn[d];
        nout *= odims[d] < 0 ? 0 : odims[d];
        stride[d] = d == 0 ? 1 : stride[d - 1] * dims[d - 1];
        if (d < N && odims[d] > 0) {
            const int64_t first = ooff[d], last = ooff[d] + (odims[d] - 1) * ostep[d];
            if (first < 0 || last > dims[d] - 1) return fail(B2F_EDIM, "requested indices exceed the image axes");
            if (border->style == B2F_INNER && (first + win_lo[d] < 0 || last + win_hi[d] > dims[d] - 1))
                return fail(B2F_EDIM, "requested indices are not in the interior for Inner()");
            if (border->style != B2F_FILL && (win_lo[d] > 0 || win_hi[d] < 0) && border->style != B2F_INNER)
                return fail(B2F_ENOTSUP, "windows that do not contain their centre need Fill or Inner borders here");
        }
    }
    if (op == 0 && wtotal > 128) return fail(B2F_ENOTSUP, "median windows hold at most 128 elements");
    if (o_numel(img) == 0) return 0;
    const bool is_i64 = img->dtype == B2F_I64;
    std::vector<double> buf(wtotal);
    std::vector<int64_t> ibuf(wtotal);
    for (int64_t o = 0; o < nout; ++o) {
        int64_t c[4], r = o;
        for (int d = 0; d < 4; ++d) { c[d] = (r % odims[d]) * ostep[d] + ooff[d]; r /= odims[d]; }
        int n = 0;
        bool nan = false;
        for (int64_t j3 = 0; j3 < wn[3]; ++j3)
            for (int64_t j2 = 0; j2 < wn[2]; ++j2)
                for (int64_t j1 = 0; j1 < wn[1]; ++j1)
                    for (int64_t j0 = 0; j0 < wn[0]; ++j0) {
                        const int64_t j[4] = {j0, j1, j2, j3};
                        int64_t lin = 0;
                        bool fillv = false;
                        for (int d = 0; d < 4; ++d) {
                            const int64_t a = c[d] + wlo[d], i = o_win_index(style, a + j[d], a, a + wn[d] - 1, dims[d]);
                            if (i < 0) fillv = true; else lin += i * stride[d];
                        }
                        if (!isf) ibuf[n] = fillv ? (int64_t)border->fill : (is_i64 ? ((const int64_t *)img->ptr)[lin] : (int64_t)o_elem(img, lin));
                        const double v = fillv ? (img->dtype == B2F_F32 ? (double)(float)border->fill : border->fill) : o_elem(img, lin);
                        nan = nan || (v != v);
                        buf[n++] = v;
                    }
        if (op != 0) {          // mean / sum / minimum / maximum in window (memory) order
            if (!isf) {
                int64_t acc = ibuf[0];
                for (int k = 1; k < n; ++k) acc = op == 3 ? std::min(acc, ibuf[k]) : op == 4 ? std::max(acc, ibuf[k]) : acc + ibuf[k];
                if (op == 1) ((double *)out->ptr)[o] = (double)acc / (double)n;
                else if (op == 2) ((int64_t *)out->ptr)[o] = acc;
                else o_store_int(out, o, acc);
            } else if (img->dtype == B2F_F32) {
                float acc = (float)buf[0];
                for (int k = 1; k < n; ++k) {
                    const float v = (float)buf[k];
                    if (op == 3) acc = v < acc ? v : acc;
                    else if (op == 4) acc = v > acc ? v : acc;
                    else { volatile float t = acc + v; acc = t; }
                }
                ((float *)out->ptr)[o] = op == 1 ? acc / (float)n : acc;
            } else {
                double acc = buf[0];
                for (int k = 1; k < n; ++k) {
                    const double v = buf[k];
                    if (op == 3) acc = v < acc ? v : acc;
                    else if (op == 4) acc = v > acc ? v : acc;
                    else { volatile double t = acc + v; acc = t; }
                }
                ((double *)out->ptr)[o] = op == 1 ? acc / (double)n : acc;
            }
            continue;
        }
        double lo_v, hi_v;
        const int mid = n / 2;
        if (is_i64) {
            std::sort(ibuf.begin(), ibuf.begin() + n);
            hi_v = (double)ibuf[mid]; lo_v = (double)ibuf[mid > 0 ? mid - 1 : 0];
        } else {
            std::sort(buf.begin(), buf.begin() + n);
            hi_v = buf[mid]; lo_v = buf[mid > 0 ? mid - 1 : 0];
        }
        if (out->dtype == B2F_F32) {
            float res;
            if (nan) res = std::numeric_limits<float>::quiet_NaN();
            else if (n & 1) res = (float)hi_v;
            else res = (float)lo_v / 2.0f + (float)hi_v / 2.0f;
            ((float *)out->ptr)[o] = res;
        } else {
            double res;
            if (nan) res = std::numeric_limits<double>::quiet_NaN();
            else if (n & 1) res = hi_v;
            else res = lo_v / 2.0 + hi_v / 2.0;
            ((double *)out->ptr)[o] = res;
        }
    }
    return 0;
}

extern "C" int b2f_mapwindow_median(const b2f_array *img, const b2f_array *out, const int64_t *win_lo, const int64_t *win_hi,
                                    const b2f_border *border, void *) {
    return o_mapwindow_reduce(img, out, 0, win_lo, win_hi, border, nullptr, nullptr);
}
extern "C" int b2f_mapwindow_reduce(const b2f_array *img, const b2f_array *out, int32_t op, const int64_t *win_lo, const int64_t *win_hi,
                                    const b2f_border *border, const int64_t *idx_first, const int64_t *idx_step, void *) {
    return o_mapwindow_reduce(img, out, op, win_lo, win_hi, border, idx_first, idx_step);
}


// ---- IIR: imfilter!(r, out, img, kernel::TriggsSdika, dim, border) (reference src/imfilter.jl:922-1092) ---------------------
// One line at a time, every operation in T = eltype(out), separate multiplies and adds in the reference's order (this file is
// compiled with -ffp-contract=off).  The 3 x 3 product kernel.M * rightΔu is StaticArrays' unrolled row . vector sum, taken here
// as ((M[r,1] d1 + M[r,2] d2) + M[r,3] d3); its association inside StaticArrays is NOT pinned at the bit level (third-party,
// not vendored) — the reference's own tests for this path (test/triggs.jl) are tolerance tests, restated in
// tests/test_iir.py.
namespace {
template <typename T>
int o_iir_typed(const b2f_array *img, const b2f_array *out, int axis, const double *coef, int style, double fill) {
    const int N = img->ndim;
    int64_t W = 1, H = img->dims[axis], B = 1;
    for (int d = 0; d < axis; ++d) W *= img->dims[d];
    for (int d = axis + 1; d < N; ++d) B *= img->dims[d];
    const T a1 = (T)coef[0], a2 = (T)coef[1], a3 = (T)coef[2], b1 = (T)coef[3], b2 = (T)coef[4], b3 = (T)coef[5], scale = (T)coef[6];
    T M[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[r][c] = (T)coef[7 + 3 * r + c];
    const T oma = (T)coef[16], omb = (T)coef[17];
    const bool copy = a1 == 0 && a2 == 0 && a3 == 0 && b1 == 0 && b2 == 0 && b3 == 0 && scale == 1;     // iscopy, :1254
    T *o = (T *)out->ptr;
    std::vector<T> x((size_t)H);
    for (int64_t bb = 0; bb < B; ++bb)
        for (int64_t w = 0; w < W; ++w) {
            const int64_t base = bb * H * W + w;
            for (int64_t i = 0; i < H; ++i) {                      // accumfilter(img[i], one(T)): the pixel in T
                int rc = load_as<T>(img->ptr, img->dtype, base + i * W, x[(size_t)i]);
                if (rc) return fail(rc, "unsupported image eltype for IIR filtering");
            }
            auto O = [&](int64_t i) -> T & { return o[base + i * W]; };
            if (copy) { for (int64_t i = 0; i < H; ++i) O(i) = x[(size_t)i]; continue; }
            const int64_t n = H;
            // leftborder!, :1017-1046
            const T iminus = style == B2F_FILL ? (T)fill : x[0];
            const T uminus = iminus / oma;
            { T t = x[0]; t += a1 * uminus; t += a2 * uminus; t += a3 * uminus; O(0) = t; }
            { T t = x[1]; t += a1 * O(0); t += a2 * uminus; t += a3 * uminus; O(1) = t; }
            { T t = x[2]; t += a1 * O(1); t += a2 * O(0); t += a3 * uminus; O(2) = t; }
            // forward, :990-998 (the last point is left to rightborder!)
            for (int64_t i = 3; i <= n - 2; ++i) { T t = x[(size_t)i]; t += a1 * O(i - 1); t += a2 * O(i - 2); t += a3 * O(i - 3); O(i) = t; }
            // rightborder!, :1048-1084
            const T iplus = style == B2F_FILL ? (T)fill : x[(size_t)(n - 1)];
            { T t = x[(size_t)(n - 1)]; t += a1 * O(n - 2); t += a2 * O(n - 3); t += a3 * O(n - 4); O(n - 1) = t; }
            const T uplus = iplus / oma, vplus = uplus / omb;
            const T d1 = O(n - 1) - uplus, d2 = O(n - 2) - uplus, d3 = O(n - 3) - uplus;
            T vr[3];
            for (int r = 0; r < 3; ++r) vr[r] = ((M[r][0] * d1 + M[r][1] * d2) + M[r][2] * d3) + vplus;
            O(n - 1) = vr[0];
            { T t = O(n - 2); t += b1 * O(n - 1); t += b2 * vr[1]; t += b3 * vr[2]; O(n - 2) = t; }
            { T t = O(n - 3); t += b1 * O(n - 2); t += b2 * O(n - 1); t += b3 * vr[1]; O(n - 3) = t; }
            // backward, :1003-1011
            for (int64_t i = n - 4; i >= 0; --i) { T t = O(i); t += b1 * O(i + 1); t += b2 * O(i + 2); t += b3 * O(i + 3); O(i) = t; }
            for (int64_t i = 0; i < n; ++i) O(i) *= scale;          // :1013-1017
        }
    return 0;
}
}  // namespace

extern "C" int b2f_iir(const b2f_array *img, const b2f_array *out, int32_t axis, const double *coef, const b2f_border *border, void *) {
    if (!img || !out || !coef || !border) return fail(B2F_EARG, "NULL argument");
    const int N = img->ndim;
    if (N < 1 || N > B2F_MAXDIM || out->ndim != N) return fail(B2F_EDIM, "IIR filtering needs 1..4 dims and equal rank");
    if (axis < 0 || axis >= N) return fail(B2F_EARG, "axis %d outside the array", (int)axis);
    for (int d = 0; d < N; ++d)
        if (img->dims[d] != out->dims[d]) return fail(B2F_EDIM, "out must have the axes of img");
    if (border->style != B2F_REPLICATE && border->style != B2F_FILL) return fail(B2F_EARG, "only \"replicate\" is supported");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_ENOTSUP, "IIR filtering produces Float32 / Float64 arrays");
    if (o_numel(img) == 0) return 0;
    bool copy = coef[6] == 1.0;
    for (int i = 0; i < 6; ++i) copy = copy && coef[i] == 0.0;
    if (!copy && img->dims[axis] <= 3)
        return fail(B2F_EDIM, "size %lld of img along dimension %d is too small for filtering with IIR kernel of length 3",
                    (long long)img->dims[axis], (int)axis + 1);
    return out->dtype == B2F_F32 ? o_iir_typed<float>(img, out, axis, coef, border->style, border->fill)
                                 : o_iir_typed<double>(img, out, axis, coef, border->style, border->fill);
}


// ---- FFT filtering (reference src/imfilter.jl:776-888) -----------------------------------------------------------------------
// The oracle has no FFT: it returns what the FFT algorithm computes up to rounding, i.e. the exact correlation with the one
// dense kernel (the reference's own tests assert `≈` between Algorithm.FIR() and Algorithm.FFT(), test/2d.jl:69-140); the GPU
// tests compare with a tolerance that covers the transforms' rounding.
extern "C" int b2f_imfilter_fft(const b2f_array *img, const b2f_array *out, const b2f_stage *kernel, const b2f_border *border,
                                const int64_t *roi_lo, const int64_t *roi_hi, void *stream) {
    if (!img || !out || !kernel || !border) return fail(B2F_EARG, "NULL argument");
    if (kernel->kind != B2F_STAGE_DENSE && kernel->kind != B2F_STAGE_1D)
        return fail(B2F_EARG, "the FFT path takes ONE array kernel (kernelconv of the factors)");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_EINEXACT, "FFT filtering produces Float32 / Float64 arrays");
    return b2f_imfilter(img, out, kernel, 1, border, roi_lo, roi_hi, stream);
}
