// kernels_pack.cu -- K1: pack + dtype-cast SoA particle fields into contiguous (N, M) chunk
// buffers, all chunks of a frame in ONE launch.
//
// Replaces, on the device, the host-side contiguity copy / cast in front of the reference's
// write: numpy.ascontiguousarray (/root/reference/pgsd/pgsd/fl.pyx:571) and
// ParticleData.validate's ascontiguousarray(dtype=float32/uint32/int32) + reshape([N,3])
// (/root/reference/pgsd/pgsd/hoomd.py:206-270):   dst[i*M + j] = (dst_type) col_j[i * stride_j].
// Casting follows numpy.astype on x86-64: float64->float32 round-to-nearest-even, integer
// narrowing wraps, NaN payloads are carried over with the quiet bit set.
//
// HBM-bound data movement: 128-bit loads (ld.global.nc.L1::no_allocate) and 128-bit stores,
// 4 rows per thread, two independent row groups in flight per thread.  A frame's chunks are
// described by a by-value segment table (<= 16 segments, 2.6 KB of kernel parameters).
#include "device_internal.h"

namespace pgsdb
{
namespace
    {
constexpr int PACK_THREADS = 256;
constexpr int ROWS_PER_THREAD = 4;
constexpr int PACK_UNROLL = 2;
constexpr int TILE_ROWS = PACK_THREADS * ROWS_PER_THREAD * PACK_UNROLL; // 2048 rows

enum PackKind : int
    {
    KIND_GENERIC = 0, // any types / strides, scalar, coalesced on dst
    KIND_W4 = 1,      // 4-byte elements moved as bits, unit strides, 16-B aligned, M <= 4
    KIND_F64_F32 = 2, // float64 columns -> float32 chunk, unit strides, aligned, M <= 4
    KIND_COPY = 3,    // M == 1 bit copy of 16-B aligned data: plain vector copy over bytes
    KIND_AOS4 = 4     // M <= 4 leading components of 16-byte records (HOOMD Scalar4 / int4 arrays), 4-byte elements
    };

struct PackSegDev
    {
    void* dst;
    const void* base[PACK_MAX_COLS];
    long long stride[PACK_MAX_COLS];
    unsigned long long N; // rows (KIND_COPY: number of 16-byte vectors, tail handled separately)
    unsigned long long tail_bytes; // KIND_COPY only
    unsigned int M;
    unsigned int tile_begin;
    unsigned char src_type, dst_type, kind, bitcopy;
    unsigned int elem_size; // dst element size
    };
struct PackArgs
    {
    PackSegDev s[PACK_MAX_SEGS];
    int nsegs;
    unsigned int total_tiles;
    };

__device__ __forceinline__ uint4 ldg_stream_v4(const void* p)
    {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
    }
__device__ __forceinline__ void stg_v4(void* p, uint4 v)
    {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
    }

// numpy/x86 semantics (cvtsd2ss): round to nearest even; NaN keeps sign and the top payload
// bits, quiet bit set.
__device__ __forceinline__ float cvt_f64_f32(double x)
    {
    float f = __double2float_rn(x);
    if (x != x)
        {
        unsigned long long b = (unsigned long long)__double_as_longlong(x);
        unsigned int r = ((unsigned int)(b >> 32) & 0x80000000u) | 0x7fc00000u
                         | (unsigned int)((b >> 29) & 0x3fffffull);
        f = __uint_as_float(r);
        }
    return f;
    }
__device__ __forceinline__ double cvt_f32_f64(float x)
    {
    double d = (double)x;
    if (x != x)
        {
        unsigned int b = __float_as_uint(x);
        unsigned long long r = ((unsigned long long)(b & 0x80000000u) << 32) | 0x7ff8000000000000ull
                               | ((unsigned long long)(b & 0x3fffffu) << 29);
        d = __longlong_as_double((long long)r);
        }
    return d;
    }

template <int M> __device__ __forceinline__ void pack_w4_rows(const PackSegDev& s, unsigned long long r)
    {
    // rows r..r+3 of every column -> 4*M consecutive words of dst
    unsigned int v[M][4];
#pragma unroll
    for (int j = 0; j < M; j++)
        {
        uint4 c = ldg_stream_v4(reinterpret_cast<const unsigned int*>(s.base[j]) + r);
        v[j][0] = c.x;
        v[j][1] = c.y;
        v[j][2] = c.z;
        v[j][3] = c.w;
        }
    unsigned int* out = reinterpret_cast<unsigned int*>(s.dst) + r * M;
#pragma unroll
    for (int q = 0; q < M; q++)
        {
        uint4 o;
        o.x = v[(4 * q + 0) % M][(4 * q + 0) / M];
        o.y = v[(4 * q + 1) % M][(4 * q + 1) / M];
        o.z = v[(4 * q + 2) % M][(4 * q + 2) / M];
        o.w = v[(4 * q + 3) % M][(4 * q + 3) / M];
        stg_v4(out + 4 * q, o);
        }
    }

// rows r..r+3 of an array of 16-byte records: component j of record i -> dst[(r+i)*M + j]
template <int M> __device__ __forceinline__ void pack_aos4_rows(const PackSegDev& s, unsigned long long r)
    {
    const uint4* in = reinterpret_cast<const uint4*>(s.base[0]) + r;
    unsigned int v[4][4]; // v[j][i]
#pragma unroll
    for (int i = 0; i < 4; i++)
        {
        uint4 a = ldg_stream_v4(in + i);
        v[0][i] = a.x;
        v[1][i] = a.y;
        v[2][i] = a.z;
        v[3][i] = a.w;
        }
    unsigned int* out = reinterpret_cast<unsigned int*>(s.dst) + r * M;
#pragma unroll
    for (int q = 0; q < M; q++)
        {
        uint4 o;
        o.x = v[(4 * q + 0) % M][(4 * q + 0) / M];
        o.y = v[(4 * q + 1) % M][(4 * q + 1) / M];
        o.z = v[(4 * q + 2) % M][(4 * q + 2) / M];
        o.w = v[(4 * q + 3) % M][(4 * q + 3) / M];
        stg_v4(out + 4 * q, o);
        }
    }

template <int M> __device__ __forceinline__ void pack_f64_rows(const PackSegDev& s, unsigned long long r)
    {
    unsigned int v[M][4];
#pragma unroll
    for (int j = 0; j < M; j++)
        {
        const double* col = reinterpret_cast<const double*>(s.base[j]) + r;
        uint4 a = ldg_stream_v4(col);
        uint4 b = ldg_stream_v4(col + 2);
        double d0 = __hiloint2double((int)a.y, (int)a.x), d1 = __hiloint2double((int)a.w, (int)a.z);
        double d2 = __hiloint2double((int)b.y, (int)b.x), d3 = __hiloint2double((int)b.w, (int)b.z);
        v[j][0] = __float_as_uint(cvt_f64_f32(d0));
        v[j][1] = __float_as_uint(cvt_f64_f32(d1));
        v[j][2] = __float_as_uint(cvt_f64_f32(d2));
        v[j][3] = __float_as_uint(cvt_f64_f32(d3));
        }
    unsigned int* out = reinterpret_cast<unsigned int*>(s.dst) + r * M;
#pragma unroll
    for (int q = 0; q < M; q++)
        {
        uint4 o;
        o.x = v[(4 * q + 0) % M][(4 * q + 0) / M];
        o.y = v[(4 * q + 1) % M][(4 * q + 1) / M];
        o.z = v[(4 * q + 2) % M][(4 * q + 2) / M];
        o.w = v[(4 * q + 3) % M][(4 * q + 3) / M];
        stg_v4(out + 4 * q, o);
        }
    }

// scalar element conversion: src element -> {signed, unsigned, floating} wide value -> dst
struct Wide
    {
    long long i;
    unsigned long long u;
    double f;
    int cls; // 0 signed, 1 unsigned, 2 float
    };

__device__ __forceinline__ Wide load_wide(const void* base, long long idx, int t)
    {
    Wide w;
    w.i = 0;
    w.u = 0;
    w.f = 0;
    w.cls = 1;
    switch (t)
        {
        case T_U8: w.u = reinterpret_cast<const uint8_t*>(base)[idx]; break;
        case T_U16: w.u = reinterpret_cast<const uint16_t*>(base)[idx]; break;
        case T_U32: w.u = reinterpret_cast<const uint32_t*>(base)[idx]; break;
        case T_U64: w.u = reinterpret_cast<const uint64_t*>(base)[idx]; break;
        case T_I8: w.i = reinterpret_cast<const int8_t*>(base)[idx]; w.cls = 0; break;
        case T_I16: w.i = reinterpret_cast<const int16_t*>(base)[idx]; w.cls = 0; break;
        case T_I32: w.i = reinterpret_cast<const int32_t*>(base)[idx]; w.cls = 0; break;
        case T_I64: w.i = reinterpret_cast<const int64_t*>(base)[idx]; w.cls = 0; break;
        case T_F32: w.f = cvt_f32_f64(reinterpret_cast<const float*>(base)[idx]); w.cls = 2; break;
        case T_F64: w.f = reinterpret_cast<const double*>(base)[idx]; w.cls = 2; break;
        default: break;
        }
    return w;
    }

// float -> integer (hoomd.py:220-266 casts whatever it is given with ascontiguousarray(dtype=uint32/int32)):
// numpy's C casts as gcc compiles them for x86-64 -- truncation toward zero.  A value whose truncation does not
// fit (or NaN) is undefined in numpy ("invalid value encountered in cast") and differs between numpy builds
// (scalar vs auto-vectorised loops); here it gives the conversion instruction's "integer indefinite" (the sign
// bit alone).  Parity is claimed and tested for values that fit.  cvttsd2si r32 serves the 8/16/32-bit signed and the 8/16-bit
// unsigned destinations, r64 serves int64; the two wide unsigned destinations go through the signed conversion
// of x - 2^(bits-1) for x >= 2^(bits-1).  float32 sources behave as their exact float64 values.
__device__ __forceinline__ unsigned int cvtt32(double d)
    {
    if (!(d < 2147483648.0) || !(d > -2147483649.0)) // NaN fails both comparisons
        return 0x80000000u;
    return (unsigned int)__double2int_rz(d);
    }
__device__ __forceinline__ unsigned long long cvtt64(double d)
    {
    if (!(d < 9223372036854775808.0) || !(d >= -9223372036854775808.0))
        return 0x8000000000000000ull;
    return (unsigned long long)__double2ll_rz(d);
    }
__device__ __forceinline__ unsigned long long float_to_int_bits(double d, int t)
    {
    switch (t)
        {
        case T_U32: return d >= 2147483648.0 ? (cvtt32(d - 2147483648.0) ^ 0x80000000u) : cvtt32(d);
        case T_I64: return cvtt64(d);
        case T_U64:
            return d >= 9223372036854775808.0 ? (cvtt64(d - 9223372036854775808.0) ^ 0x8000000000000000ull) : cvtt64(d);
        default: return cvtt32(d); // narrowed by the store
        }
    }

__device__ __forceinline__ void store_wide(void* dst, unsigned long long e, int t, const Wide& w)
    {
    // integer destinations: modular narrowing of the 64-bit pattern (numpy astype wraps)
    unsigned long long bits = w.cls == 0 ? (unsigned long long)w.i : w.u;
    if (w.cls == 2 && t != T_F32 && t != T_F64)
        bits = float_to_int_bits(w.f, t);
    switch (t)
        {
        case T_U8:
        case T_I8: reinterpret_cast<uint8_t*>(dst)[e] = (uint8_t)bits; break;
        case T_U16:
        case T_I16: reinterpret_cast<uint16_t*>(dst)[e] = (uint16_t)bits; break;
        case T_U32:
        case T_I32: reinterpret_cast<uint32_t*>(dst)[e] = (uint32_t)bits; break;
        case T_U64:
        case T_I64: reinterpret_cast<uint64_t*>(dst)[e] = bits; break;
        case T_F32:
            {
            float f = w.cls == 2 ? cvt_f64_f32(w.f) : (w.cls == 0 ? __ll2float_rn(w.i) : __ull2float_rn(w.u));
            reinterpret_cast<float*>(dst)[e] = f;
            break;
            }
        case T_F64:
            {
            double d = w.cls == 2 ? w.f : (w.cls == 0 ? __ll2double_rn(w.i) : __ull2double_rn(w.u));
            reinterpret_cast<double*>(dst)[e] = d;
            break;
            }
        default: break;
        }
    }

__device__ __forceinline__ void pack_generic_elem(const PackSegDev& s, unsigned long long row, unsigned int col)
    {
    const unsigned long long e = row * s.M + col;
    const long long idx = (long long)row * s.stride[col];
    if (s.bitcopy)
        {
        switch (s.elem_size)
            {
            case 1: reinterpret_cast<uint8_t*>(s.dst)[e] = reinterpret_cast<const uint8_t*>(s.base[col])[idx]; break;
            case 2: reinterpret_cast<uint16_t*>(s.dst)[e] = reinterpret_cast<const uint16_t*>(s.base[col])[idx]; break;
            case 4: reinterpret_cast<uint32_t*>(s.dst)[e] = reinterpret_cast<const uint32_t*>(s.base[col])[idx]; break;
            default: reinterpret_cast<uint64_t*>(s.dst)[e] = reinterpret_cast<const uint64_t*>(s.base[col])[idx]; break;
            }
        }
    else
        store_wide(s.dst, e, s.dst_type, load_wide(s.base[col], idx, s.src_type));
    }

__global__ void __launch_bounds__(PACK_THREADS) k1_pack_frame(const __grid_constant__ PackArgs args)
    {
    for (unsigned int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x)
        {
        int si = 0;
#pragma unroll 1
        for (int k = 1; k < args.nsegs; k++)
            if (tile >= args.s[k].tile_begin)
                si = k;
        const PackSegDev& s = args.s[si];
        const unsigned long long row0 = (unsigned long long)(tile - s.tile_begin) * TILE_ROWS;
        if (s.kind == KIND_W4 || s.kind == KIND_F64_F32 || s.kind == KIND_AOS4)
            {
#pragma unroll
            for (int it = 0; it < PACK_UNROLL; it++)
                {
                unsigned long long r = row0 + (unsigned long long)it * (PACK_THREADS * ROWS_PER_THREAD)
                                       + (unsigned long long)threadIdx.x * ROWS_PER_THREAD;
                if (r + ROWS_PER_THREAD <= s.N)
                    {
                    if (s.kind == KIND_AOS4)
                        {
                        switch (s.M)
                            {
                            case 1: pack_aos4_rows<1>(s, r); break;
                            case 2: pack_aos4_rows<2>(s, r); break;
                            case 3: pack_aos4_rows<3>(s, r); break;
                            default: pack_aos4_rows<4>(s, r); break;
                            }
                        }
                    else if (s.kind == KIND_W4)
                        {
                        switch (s.M)
                            {
                            case 1: pack_w4_rows<1>(s, r); break;
                            case 2: pack_w4_rows<2>(s, r); break;
                            case 3: pack_w4_rows<3>(s, r); break;
                            default: pack_w4_rows<4>(s, r); break;
                            }
                        }
                    else
                        {
                        switch (s.M)
                            {
                            case 1: pack_f64_rows<1>(s, r); break;
                            case 2: pack_f64_rows<2>(s, r); break;
                            case 3: pack_f64_rows<3>(s, r); break;
                            default: pack_f64_rows<4>(s, r); break;
                            }
                        }
                    }
                else
                    {
                    for (unsigned long long rr = r; rr < s.N && rr < r + ROWS_PER_THREAD; rr++)
                        for (unsigned int c = 0; c < s.M; c++)
                            pack_generic_elem(s, rr, c);
                    }
                }
            }
        else if (s.kind == KIND_COPY)
            {
            const uint4* in = reinterpret_cast<const uint4*>(s.base[0]);
            uint4* out = reinterpret_cast<uint4*>(s.dst);
            uint4 v[ROWS_PER_THREAD * PACK_UNROLL];
#pragma unroll
            for (int k = 0; k < ROWS_PER_THREAD * PACK_UNROLL; k++)
                {
                unsigned long long i = row0 + (unsigned long long)k * PACK_THREADS + threadIdx.x;
                if (i < s.N)
                    v[k] = ldg_stream_v4(in + i);
                }
#pragma unroll
            for (int k = 0; k < ROWS_PER_THREAD * PACK_UNROLL; k++)
                {
                unsigned long long i = row0 + (unsigned long long)k * PACK_THREADS + threadIdx.x;
                if (i < s.N)
                    stg_v4(out + i, v[k]);
                }
            if (row0 == 0 && threadIdx.x < s.tail_bytes)
                reinterpret_cast<unsigned char*>(s.dst)[s.N * 16 + threadIdx.x]
                    = reinterpret_cast<const unsigned char*>(s.base[0])[s.N * 16 + threadIdx.x];
            }
        else
            {
            unsigned long long rows = s.N - row0 < (unsigned long long)TILE_ROWS ? s.N - row0 : TILE_ROWS;
            unsigned long long elems = rows * s.M;
            for (unsigned long long q = threadIdx.x; q < elems; q += PACK_THREADS)
                {
                unsigned long long row = q / s.M;
                unsigned int col = (unsigned int)(q - row * s.M);
                pack_generic_elem(s, row0 + row, col);
                }
            }
        }
    }
    } // namespace

size_t type_size(int t)
    {
    switch (t)
        {
        case T_U8:
        case T_I8: return 1;
        case T_U16:
        case T_I16: return 2;
        case T_U32:
        case T_I32:
        case T_F32: return 4;
        case T_U64:
        case T_I64:
        case T_F64: return 8;
        default: return 0;
        }
    }

static inline bool is_float_type(int t) { return t == T_F32 || t == T_F64; }

bool cast_supported(int src, int dst)
    {
    if (type_size(src) == 0 || type_size(dst) == 0)
        return false;
    return true;
    }

int pack_launch(const PackSegment* segs, int nsegs, cudaStream_t st)
    {
    if (nsegs <= 0)
        return 0;
    if (nsegs > PACK_MAX_SEGS)
        {
        set_last_error("pack: too many segments in one launch");
        return -2;
        }
    PackArgs a;
    memset(&a, 0, sizeof(a));
    unsigned long long tiles = 0;
    int k = 0;
    for (int i = 0; i < nsegs; i++)
        {
        const PackSegment& g = segs[i];
        if (g.M == 0 || g.M > (unsigned)PACK_MAX_COLS || !cast_supported(g.src_type, g.dst_type))
            {
            set_last_error("pack: unsupported segment (M must be 1..8, known element types)");
            return -2;
            }
        if (g.N == 0)
            continue;
        PackSegDev& s = a.s[k];
        s.dst = g.dst;
        s.M = g.M;
        s.src_type = (unsigned char)g.src_type;
        s.dst_type = (unsigned char)g.dst_type;
        s.elem_size = (unsigned int)type_size(g.dst_type);
        const bool same_size = type_size(g.src_type) == type_size(g.dst_type);
        s.bitcopy = (same_size && (g.src_type == g.dst_type || (!is_float_type(g.src_type) && !is_float_type(g.dst_type))))
                        ? 1
                        : 0;
        bool unit = true, aligned = ((uintptr_t)g.dst % 16 == 0);
        for (unsigned j = 0; j < g.M; j++)
            {
            if (g.base[j] == nullptr)
                {
                set_last_error("pack: NULL column");
                return -2;
                }
            s.base[j] = g.base[j];
            s.stride[j] = g.stride[j];
            unit = unit && (g.stride[j] == 1);
            aligned = aligned && ((uintptr_t)g.base[j] % 16 == 0);
            }
        s.N = g.N;
        s.kind = KIND_GENERIC;
        unsigned long long ntiles = (g.N + TILE_ROWS - 1) / TILE_ROWS;
        if (unit && aligned && g.M == 1 && s.bitcopy)
            {
            unsigned long long bytes = g.N * s.elem_size;
            s.kind = KIND_COPY;
            s.N = bytes / 16;
            s.tail_bytes = bytes % 16;
            ntiles = (s.N + TILE_ROWS - 1) / TILE_ROWS;
            if (ntiles == 0)
                ntiles = 1;
            }
        else if (unit && aligned && g.M <= 4 && s.bitcopy && s.elem_size == 4)
            s.kind = KIND_W4;
        else if (unit && aligned && g.M <= 4 && g.src_type == T_F64 && g.dst_type == T_F32)
            s.kind = KIND_F64_F32;
        else if (g.M <= 4 && s.bitcopy && s.elem_size == 4 && (uintptr_t)g.dst % 16 == 0 && (uintptr_t)g.base[0] % 16 == 0)
            {
            // components 0..M-1 of 16-byte records: every column is base[0] + 4*j with stride 4
            bool rec = true;
            for (unsigned j = 0; j < g.M; j++)
                rec = rec && g.stride[j] == 4 && (const char*)g.base[j] == (const char*)g.base[0] + 4 * j;
            if (rec)
                s.kind = KIND_AOS4;
            }
        if (tiles + ntiles > 0xffffffffull)
            {
            set_last_error("pack: frame too large for one launch");
            return -2;
            }
        s.tile_begin = (unsigned int)tiles;
        tiles += ntiles;
        k++;
        }
    if (k == 0)
        return 0;
    a.nsegs = k;
    a.total_tiles = (unsigned int)tiles;
    unsigned long long grid = tiles;
    unsigned long long cap = (unsigned long long)dev_sm_count() * 8; // 8 resident CTAs of 256 threads per SM
    if (grid > cap)
        grid = cap;
    k1_pack_frame<<<(unsigned)grid, PACK_THREADS, 0, st>>>(a);
    dev_stats().kernel_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        {
        set_last_error(std::string("pack launch: ") + cudaGetErrorString(e));
        return -1;
        }
    return 0;
    }
} // namespace pgsdb
