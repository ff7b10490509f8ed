// oracle/clc_compat.hpp -- TEST INFRASTRUCTURE, not product code.
//
// A small OpenCL-C-1.0-in-C++ compatibility layer.  Its only purpose is to let
// g++ compile the reference's own kernel files (/root/reference/src/GPU_kernels.cl
// and CPU_kernels.cl) for the host CPU, unmodified, so that the parity oracle can
// be pinned against the REAL reference arithmetic (this image has no OpenCL
// compiler or runtime).  It provides exactly the
// language features those two files use: address-space qualifiers, the vector
// types with their .xyzw / .sN / .sNNNN swizzles, element-wise operators with
// OpenCL semantics (vector compares yield -1/0, integer lanes wrap at the lane
// width, scalars widen to the vector's element type), convert_*[_sat], vload/
// vstore, select, abs, mad, mad24, read_imageui with a clamp-to-edge nearest
// sampler, and the work-item id functions.
//
// Nothing in vp8oclenc_b200/ includes this file.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <utility>

namespace clc {

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;

// ---------------------------------------------------------------- work-item state
struct work_item_state {
    size_t global_id, local_id, local_size, global_size, group_id;
};
extern thread_local work_item_state g_wi;
inline size_t get_global_id(uint) { return g_wi.global_id; }
inline size_t get_local_id(uint) { return g_wi.local_id; }
inline size_t get_local_size(uint) { return g_wi.local_size; }
inline size_t get_global_size(uint) { return g_wi.global_size; }
inline size_t get_group_id(uint) { return g_wi.group_id; }

// ---------------------------------------------------------------- vector machinery
template <class D, class T, int M>
struct vx {  // CRTP tag: "something that reads/writes like an M-vector of T"
    const T &at(int i) const { return static_cast<const D *>(this)->ref(i); }
    T &at(int i) { return static_cast<D *>(this)->ref(i); }
};

template <class T, int N>
struct vec;

// swizzle proxy: elements [OFF, OFF+M) of an N-vector that it aliases inside a union
template <class T, int N, int OFF, int M>
struct sub : vx<sub<T, N, OFF, M>, T, M> {
    T d[N];
    sub() = default;
    const T &ref(int i) const { return d[OFF + i]; }
    T &ref(int i) { return d[OFF + i]; }
    sub &operator=(const sub &o) {
        T t[M];
        for (int i = 0; i < M; ++i) t[i] = o.d[OFF + i];
        for (int i = 0; i < M; ++i) d[OFF + i] = t[i];
        return *this;
    }
    template <class E>
    sub &operator=(const vx<E, T, M> &o) {
        T t[M];
        for (int i = 0; i < M; ++i) t[i] = o.at(i);
        for (int i = 0; i < M; ++i) d[OFF + i] = t[i];
        return *this;
    }
    operator vec<T, M>() const;
};

#define CLC_VEC_COMMON(N)                                                                \
    vec() {}                                                                             \
    vec(const vec &o) { std::memcpy(s, o.s, sizeof(s)); }                                \
    vec &operator=(const vec &o) {                                                       \
        std::memmove(s, o.s, sizeof(s));                                                 \
        return *this;                                                                    \
    }                                                                                    \
    vec(T v) {                                                                           \
        for (int i = 0; i < N; ++i) s[i] = v;                                            \
    }                                                                                    \
    template <class E>                                                                   \
    vec(const vx<E, T, N> &o) {                                                          \
        for (int i = 0; i < N; ++i) s[i] = o.at(i);                                      \
    }                                                                                    \
    template <class... A, class = typename std::enable_if<sizeof...(A) == N>::type>     \
    vec(A... a) : s{(T)a...} {}                                                          \
    const T &ref(int i) const { return s[i]; }                                           \
    T &ref(int i) { return s[i]; }

template <class T>
struct vec<T, 2> : vx<vec<T, 2>, T, 2> {
    union {
        T s[2];
        struct { T x, y; };
        struct { T s0, s1; };
    };
    CLC_VEC_COMMON(2)
};

template <class T>
struct vec<T, 4> : vx<vec<T, 4>, T, 4> {
    union {
        T s[4];
        struct { T x, y, z, w; };
        struct { T s0, s1, s2, s3; };
        sub<T, 4, 0, 2> s01, xy, lo;
        sub<T, 4, 2, 2> s23, zw, hi;
    };
    CLC_VEC_COMMON(4)
};

template <class T>
struct vec<T, 8> : vx<vec<T, 8>, T, 8> {
    union {
        T s[8];
        struct { T s0, s1, s2, s3, s4, s5, s6, s7; };
        sub<T, 8, 0, 2> s01;
        sub<T, 8, 2, 2> s23;
        sub<T, 8, 4, 2> s45;
        sub<T, 8, 6, 2> s67;
        sub<T, 8, 0, 4> s0123, lo;
        sub<T, 8, 4, 4> s4567, hi;
    };
    CLC_VEC_COMMON(8)
};

template <class T>
struct vec<T, 16> : vx<vec<T, 16>, T, 16> {
    union {
        T s[16];
        struct { T s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, sA, sB, sC, sD, sE, sF; };
        sub<T, 16, 0, 2> s01;
        sub<T, 16, 2, 2> s23;
        sub<T, 16, 4, 2> s45;
        sub<T, 16, 6, 2> s67;
        sub<T, 16, 8, 2> s89;
        sub<T, 16, 10, 2> sAB;
        sub<T, 16, 12, 2> sCD;
        sub<T, 16, 14, 2> sEF;
        sub<T, 16, 0, 4> s0123;
        sub<T, 16, 4, 4> s4567;
        sub<T, 16, 8, 4> s89AB;
        sub<T, 16, 12, 4> sCDEF;
        sub<T, 16, 0, 8> lo;
        sub<T, 16, 8, 8> hi;
    };
    CLC_VEC_COMMON(16)
};

template <class T, int N, int OFF, int M>
sub<T, N, OFF, M>::operator vec<T, M>() const {
    vec<T, M> r;
    for (int i = 0; i < M; ++i) r.s[i] = d[OFF + i];
    return r;
}

#define CLC_TYPES(T)              \
    typedef vec<T, 2> T##2;       \
    typedef vec<T, 4> T##4;       \
    typedef vec<T, 8> T##8;       \
    typedef vec<T, 16> T##16;
typedef signed char schar;
CLC_TYPES(uchar)
CLC_TYPES(short)
CLC_TYPES(ushort)
CLC_TYPES(int)
CLC_TYPES(uint)
CLC_TYPES(float)
typedef vec<schar, 2> char2;
typedef vec<schar, 4> char4;
typedef vec<schar, 8> char8;
typedef vec<schar, 16> char16;
static_assert(sizeof(uchar4) == 4 && sizeof(short2) == 4 && sizeof(int4) == 16 &&
                  sizeof(int16) == 64 && sizeof(uchar16) == 16 && sizeof(short8) == 16 &&
                  sizeof(float4) == 16,
              "vector layouts must match OpenCL's");

// result type of a vector comparison: signed integer lanes of the same width
template <class T> struct cmp_of { typedef typename std::make_signed<T>::type type; };
template <> struct cmp_of<float> { typedef int type; };

template <class S> struct is_scalar {
    static const bool value = std::is_arithmetic<S>::value || std::is_enum<S>::value;
};
#define CLC_IF_SCALAR(S) typename std::enable_if<is_scalar<S>::value, int>::type = 0

// element-wise binary operators; the result lane is computed in C's promoted type and
// truncated back to T, which is what an OpenCL device does for 8/16-bit lanes
#define CLC_BINOP(OP)                                                                       \
    template <class A, class B, class T, int M>                                             \
    vec<T, M> operator OP(const vx<A, T, M> &a, const vx<B, T, M> &b) {                     \
        vec<T, M> r;                                                                        \
        for (int i = 0; i < M; ++i) r.s[i] = (T)(a.at(i) OP b.at(i));                       \
        return r;                                                                           \
    }                                                                                       \
    template <class A, class T, int M, class S, CLC_IF_SCALAR(S)>                           \
    vec<T, M> operator OP(const vx<A, T, M> &a, S b) {                                      \
        vec<T, M> r;                                                                        \
        for (int i = 0; i < M; ++i) r.s[i] = (T)(a.at(i) OP(T) b);                          \
        return r;                                                                           \
    }                                                                                       \
    template <class A, class T, int M, class S, CLC_IF_SCALAR(S)>                           \
    vec<T, M> operator OP(S a, const vx<A, T, M> &b) {                                      \
        vec<T, M> r;                                                                        \
        for (int i = 0; i < M; ++i) r.s[i] = (T)((T)a OP b.at(i));                          \
        return r;                                                                           \
    }                                                                                       \
    template <class A, class B, class T, int M>                                             \
    A &operator OP##=(vx<A, T, M> &a, const vx<B, T, M> &b) {                               \
        T t[M];                                                                             \
        for (int i = 0; i < M; ++i) t[i] = (T)(a.at(i) OP b.at(i));                         \
        for (int i = 0; i < M; ++i) a.at(i) = t[i];                                         \
        return static_cast<A &>(a);                                                         \
    }                                                                                       \
    template <class A, class T, int M, class S, CLC_IF_SCALAR(S)>                           \
    A &operator OP##=(vx<A, T, M> &a, S b) {                                                \
        for (int i = 0; i < M; ++i) a.at(i) = (T)(a.at(i) OP(T) b);                         \
        return static_cast<A &>(a);                                                         \
    }
CLC_BINOP(+)
CLC_BINOP(-)
CLC_BINOP(*)
CLC_BINOP(/)
CLC_BINOP(%)
CLC_BINOP(&)
CLC_BINOP(|)
CLC_BINOP(^)

// shifts: the count is a plain scalar or a vector; arithmetic for signed lanes
#define CLC_SHIFT(OP)                                                                       \
    template <class A, class T, int M, class S, CLC_IF_SCALAR(S)>                           \
    vec<T, M> operator OP(const vx<A, T, M> &a, S n) {                                      \
        vec<T, M> r;                                                                        \
        for (int i = 0; i < M; ++i) r.s[i] = (T)(a.at(i) OP(int) n);                        \
        return r;                                                                           \
    }                                                                                       \
    template <class A, class T, int M, class S, CLC_IF_SCALAR(S)>                           \
    A &operator OP##=(vx<A, T, M> &a, S n) {                                                \
        for (int i = 0; i < M; ++i) a.at(i) = (T)(a.at(i) OP(int) n);                       \
        return static_cast<A &>(a);                                                         \
    }
CLC_SHIFT(<<)
CLC_SHIFT(>>)

#define CLC_CMP(OP)                                                                         \
    template <class A, class B, class T, int M>                                             \
    vec<typename cmp_of<T>::type, M> operator OP(const vx<A, T, M> &a, const vx<B, T, M> &b) { \
        vec<typename cmp_of<T>::type, M> r;                                                 \
        for (int i = 0; i < M; ++i) r.s[i] = (a.at(i) OP b.at(i)) ? -1 : 0;                 \
        return r;                                                                           \
    }                                                                                       \
    template <class A, class T, int M, class S, CLC_IF_SCALAR(S)>                           \
    vec<typename cmp_of<T>::type, M> operator OP(const vx<A, T, M> &a, S b) {               \
        vec<typename cmp_of<T>::type, M> r;                                                 \
        for (int i = 0; i < M; ++i) r.s[i] = (a.at(i) OP(T) b) ? -1 : 0;                    \
        return r;                                                                           \
    }
CLC_CMP(<)
CLC_CMP(>)
CLC_CMP(<=)
CLC_CMP(>=)
CLC_CMP(==)
CLC_CMP(!=)

template <class A, class T, int M>
vec<T, M> operator~(const vx<A, T, M> &a) {
    vec<T, M> r;
    for (int i = 0; i < M; ++i) r.s[i] = (T)(~a.at(i));
    return r;
}
template <class A, class T, int M>
vec<T, M> operator-(const vx<A, T, M> &a) {
    vec<T, M> r;
    for (int i = 0; i < M; ++i) r.s[i] = (T)(-a.at(i));
    return r;
}

// ---------------------------------------------------------------- conversions
template <class D, class S>
inline D sat_cast(S v) {
    if (std::is_floating_point<S>::value) {
        double lo = (double)std::numeric_limits<D>::min(), hi = (double)std::numeric_limits<D>::max();
        double x = (double)v;
        return (D)(x < lo ? lo : (x > hi ? hi : x));
    }
    long long lo = (long long)std::numeric_limits<D>::min(), hi = (long long)std::numeric_limits<D>::max();
    long long x = (long long)v;
    return (D)(x < lo ? lo : (x > hi ? hi : x));
}
}  // namespace clc
#include <limits>
namespace clc {

#define CLC_CONVERT(D, NAME, N)                                                             \
    template <class A, class S>                                                             \
    vec<D, N> convert_##NAME##N(const vx<A, S, N> &a) {                                     \
        vec<D, N> r;                                                                        \
        for (int i = 0; i < N; ++i) r.s[i] = (D)a.at(i);                                    \
        return r;                                                                           \
    }                                                                                       \
    template <class A, class S>                                                             \
    vec<D, N> convert_##NAME##N##_sat(const vx<A, S, N> &a) {                               \
        vec<D, N> r;                                                                        \
        for (int i = 0; i < N; ++i) r.s[i] = sat_cast<D, S>(a.at(i));                       \
        return r;                                                                           \
    }
#define CLC_CONVERT_ALL(D, NAME) \
    CLC_CONVERT(D, NAME, 2) CLC_CONVERT(D, NAME, 4) CLC_CONVERT(D, NAME, 8) CLC_CONVERT(D, NAME, 16)
CLC_CONVERT_ALL(uchar, uchar)
CLC_CONVERT_ALL(schar, char)
CLC_CONVERT_ALL(short, short)
CLC_CONVERT_ALL(ushort, ushort)
CLC_CONVERT_ALL(int, int)
CLC_CONVERT_ALL(uint, uint)
CLC_CONVERT_ALL(float, float)

// ---------------------------------------------------------------- vload / vstore
#define CLC_VLOADSTORE(N)                                                                   \
    template <class T>                                                                      \
    vec<typename std::remove_cv<T>::type, N> vload##N(size_t off, T *p) {                   \
        vec<typename std::remove_cv<T>::type, N> r;                                         \
        std::memcpy(r.s, p + off * N, sizeof(r.s));                                         \
        return r;                                                                           \
    }                                                                                       \
    template <class A, class T>                                                             \
    void vstore##N(const vx<A, T, N> &v, size_t off, T *p) {                                \
        for (int i = 0; i < N; ++i) p[off * N + i] = v.at(i);                               \
    }
CLC_VLOADSTORE(2)
CLC_VLOADSTORE(4)
CLC_VLOADSTORE(8)
CLC_VLOADSTORE(16)

// ---------------------------------------------------------------- built-ins
// abs() of a signed integer returns the unsigned type (OpenCL 1.0 6.11.3)
inline uint abs(int v) { return v < 0 ? 0u - (uint)v : (uint)v; }
inline uint abs(uint v) { return v; }
template <class A, class T, int M>
vec<typename std::make_unsigned<T>::type, M> abs(const vx<A, T, M> &a) {
    typedef typename std::make_unsigned<T>::type U;
    vec<U, M> r;
    for (int i = 0; i < M; ++i) {
        T v = a.at(i);
        r.s[i] = v < 0 ? (U)(0 - (U)v) : (U)v;
    }
    return r;
}

inline int mad24(int a, int b, int c) { return a * b + c; }
template <class A, class B, int M, class S, CLC_IF_SCALAR(S)>
vec<int, M> mad24(const vx<A, int, M> &a, S b, const vx<B, int, M> &c) {
    vec<int, M> r;
    for (int i = 0; i < M; ++i) r.s[i] = a.at(i) * (int)b + c.at(i);
    return r;
}

// mad(): the parity definition fixed in SURVEY.md Q10 is a fused multiply-add
inline float mad(float a, float b, float c) { return fmaf(a, b, c); }
template <class A, class B, class C, int M>
vec<float, M> mad(const vx<A, float, M> &a, const vx<B, float, M> &b, const vx<C, float, M> &c) {
    vec<float, M> r;
    for (int i = 0; i < M; ++i) r.s[i] = fmaf(a.at(i), b.at(i), c.at(i));
    return r;
}

// scalar select(a, b, c) = c ? b : a ; vector select picks b where c's MSB is set
template <class A, class B, class C,
          typename std::enable_if<is_scalar<A>::value && is_scalar<B>::value && is_scalar<C>::value, int>::type = 0>
typename std::decay<decltype(true ? std::declval<B>() : std::declval<A>())>::type select(A a, B b, C c) {
    return c ? b : a;
}
template <class A, class B, class C, class T, class CT, int M>
vec<T, M> select(const vx<A, T, M> &a, const vx<B, T, M> &b, const vx<C, CT, M> &c) {
    vec<T, M> r;
    for (int i = 0; i < M; ++i) r.s[i] = (c.at(i) < 0) ? b.at(i) : a.at(i);
    return r;
}
template <class A, class C, class T, class CT, int M, class S, CLC_IF_SCALAR(S)>
vec<T, M> select(const vx<A, T, M> &a, S b, const vx<C, CT, M> &c) {
    vec<T, M> r;
    for (int i = 0; i < M; ++i) r.s[i] = (c.at(i) < 0) ? (T)b : a.at(i);
    return r;
}

// ---------------------------------------------------------------- images
struct image2d {
    const uchar *data;  // CL_R / CL_UNSIGNED_INT8, tightly packed rows
    int width, height;
};
typedef const image2d *image2d_t;
typedef int sampler_t;
enum { CLK_NORMALIZED_COORDS_FALSE = 0, CLK_ADDRESS_CLAMP_TO_EDGE = 2, CLK_FILTER_NEAREST = 0x10 };
inline uint4 read_imageui(image2d_t img, sampler_t, const int2 &c) {
    int x = c.x < 0 ? 0 : (c.x >= img->width ? img->width - 1 : c.x);
    int y = c.y < 0 ? 0 : (c.y >= img->height ? img->height - 1 : c.y);
    return uint4((uint)img->data[(size_t)y * img->width + x], 0u, 0u, 1u);
}

// ---------------------------------------------------------------- kernel registry
// One descriptor per __kernel; invoke() receives, per argument, a pointer to the value
// (scalars: the bytes given to clSetKernelArg; pointers and images: a host pointer the
// runtime resolved from the cl_mem handle).
struct kernel_desc {
    const char *name;
    int nargs;
    const unsigned char *is_pointer;  // nargs flags
    void (*invoke)(void *const *argv);
};

template <class... A, size_t... I>
inline void call_with(void (*fn)(A...), void *const *argv, std::index_sequence<I...>) {
    fn(*(typename std::remove_cv<typename std::remove_reference<A>::type>::type *)argv[I]...);
}
template <class... A>
struct arg_kinds {
    static const unsigned char value[sizeof...(A) + 1];
};
template <class... A>
const unsigned char arg_kinds<A...>::value[sizeof...(A) + 1] = {(unsigned char)std::is_pointer<A>::value..., 0};
template <class... A>
inline int count_args(void (*)(A...)) { return (int)sizeof...(A); }
template <class... A>
inline const unsigned char *kinds_of(void (*)(A...)) { return arg_kinds<A...>::value; }

#define CLC_KERNEL_ENTRY(NS, NAME)                                                          \
    {#NAME, clc::count_args(&NS::NAME), clc::kinds_of(&NS::NAME), [](void *const *argv) {   \
         clc::invoke_fn(&NS::NAME, argv);                                                   \
     }}
template <class... A>
inline void invoke_fn(void (*fn)(A...), void *const *argv) {
    call_with(fn, argv, std::index_sequence_for<A...>{});
}

}  // namespace clc

// ---------------------------------------------------------------- qualifiers
#define __kernel
#define __global
#define __local
#define __constant const
#define __private
#define __read_only
#define __write_only
#define restrict __restrict__
