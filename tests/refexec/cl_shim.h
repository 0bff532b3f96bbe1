// OpenCL C -> C++ shim used by tests/refexec to execute the reference's kernel sources (rendered
// from the templates that lie under /root/reference) on the CPU with g++.  Test infrastructure.
//
// Only what those kernels use: address-space qualifiers, the vector types double2/3/4 and
// float2/3/4 with .x/.y/.z/.w and .s0-.s3 members and component-wise arithmetic, fmin/fmax/fabs,
// the integer builtins, and the work-item functions of a serial "one work item at a time" run.
// Build with -ffp-contract=off so a*b+c is never fused (the CUDA side uses -fmad=false).
#pragma once
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define __global
#define __local
#define __constant const
#define __private
#define __kernel
#define global
#define restrict __restrict__
#define barrier(x)
#define CLK_LOCAL_MEM_FENCE 0
#define CLK_GLOBAL_MEM_FENCE 0

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;

template <class T, int N> struct clvec;

template <class T> struct clvec<T, 2> {
    union { struct { T x, y; }; struct { T s0, s1; }; T v[2]; };
    clvec() {}
    clvec(T a) : x(a), y(a) {}
    clvec(T a, T b) : x(a), y(b) {}
};
template <class T> struct clvec<T, 3> {
    union { struct { T x, y, z; }; struct { T s0, s1, s2; }; T v[4]; };
    clvec() {}
    clvec(T a) : x(a), y(a), z(a) {}
    clvec(T a, T b, T c) : x(a), y(b), z(c) {}
};
template <class T> struct clvec<T, 4> {
    union { struct { T x, y, z, w; }; struct { T s0, s1, s2, s3; }; T v[4]; };
    clvec() {}
    clvec(T a) : x(a), y(a), z(a), w(a) {}
    clvec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
};

#define CLVEC_BINOP(op)                                                                  \
    template <class T, int N> inline clvec<T, N> operator op(const clvec<T, N> &a,       \
                                                             const clvec<T, N> &b) {     \
        clvec<T, N> r;                                                                   \
        for (int i = 0; i < N; ++i) r.v[i] = a.v[i] op b.v[i];                           \
        return r;                                                                        \
    }                                                                                    \
    template <class T, int N> inline clvec<T, N> operator op(const clvec<T, N> &a, T b) { \
        clvec<T, N> r;                                                                   \
        for (int i = 0; i < N; ++i) r.v[i] = a.v[i] op b;                                \
        return r;                                                                        \
    }                                                                                    \
    template <class T, int N> inline clvec<T, N> operator op(T a, const clvec<T, N> &b) { \
        clvec<T, N> r;                                                                   \
        for (int i = 0; i < N; ++i) r.v[i] = a op b.v[i];                                \
        return r;                                                                        \
    }
CLVEC_BINOP(+)
CLVEC_BINOP(-)
CLVEC_BINOP(*)
CLVEC_BINOP(/)
#undef CLVEC_BINOP

template <class T, int N> inline clvec<T, N> fabs(const clvec<T, N> &a) {
    clvec<T, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = std::fabs(a.v[i]);
    return r;
}
template <class T, int N> inline clvec<T, N> fmin(const clvec<T, N> &a, const clvec<T, N> &b) {
    clvec<T, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = std::fmin(a.v[i], b.v[i]);
    return r;
}
template <class T, int N> inline clvec<T, N> fmax(const clvec<T, N> &a, const clvec<T, N> &b) {
    clvec<T, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = std::fmax(a.v[i], b.v[i]);
    return r;
}

typedef clvec<double, 2> double2;
typedef clvec<double, 3> double3;
typedef clvec<double, 4> double4;
typedef clvec<float, 2> float2;
typedef clvec<float, 3> float3;
typedef clvec<float, 4> float4;
typedef clvec<int, 2> int2;
typedef clvec<int, 3> int3;
typedef clvec<int, 4> int4;

// scalar builtins: keep the argument type (OpenCL's fmin(float, float) is a float operation)
inline float fmin(float a, float b) { return std::fmin(a, b); }
inline float fmax(float a, float b) { return std::fmax(a, b); }
inline float fabs(float a) { return std::fabs(a); }
inline double fmin(double a, double b) { return std::fmin(a, b); }
inline double fmax(double a, double b) { return std::fmax(a, b); }
inline double fabs(double a) { return std::fabs(a); }
// <cmath> leaves only the C (double) versions of these in the global namespace: without the
// overloads a float argument would silently be evaluated in double, unlike OpenCL C
inline float sqrt(float a) { return std::sqrt(a); }
inline float rint(float a) { return std::rint(a); }
inline float floor(float a) { return std::floor(a); }
inline float ceil(float a) { return std::ceil(a); }
// OpenCL: min(x, y) is y if y < x, otherwise x; max(x, y) is y if x < y, otherwise x
template <class T> inline T clmin(T a, T b) { return b < a ? b : a; }
template <class T> inline T clmax(T a, T b) { return a < b ? b : a; }
template <class T, int N> inline clvec<T, N> clmin(const clvec<T, N> &a, const clvec<T, N> &b) {
    clvec<T, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = b.v[i] < a.v[i] ? b.v[i] : a.v[i];
    return r;
}
template <class T, int N> inline clvec<T, N> clmax(const clvec<T, N> &a, const clvec<T, N> &b) {
    clvec<T, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = a.v[i] < b.v[i] ? b.v[i] : a.v[i];
    return r;
}
#define min(a, b) clmin(a, b)
#define max(a, b) clmax(a, b)

// geometric builtins on scalars: distance(a, b) = length(a - b) = |a - b|  (sqrt(x*x) == |x| in
// binary IEEE arithmetic barring over/underflow, so either formulation gives the same bits)
inline float distance(float a, float b) { return std::fabs(a - b); }
inline double distance(double a, double b) { return std::fabs(a - b); }

// saturating integer add (OpenCL add_sat)
inline int add_sat(int a, int b) {
    long r = (long) a + (long) b;
    return r > INT_MAX ? INT_MAX : (r < INT_MIN ? INT_MIN : (int) r);
}
inline unsigned int add_sat(unsigned int a, unsigned int b) {
    unsigned long r = (unsigned long) a + (unsigned long) b;
    return r > UINT_MAX ? UINT_MAX : (unsigned int) r;
}
inline long add_sat(long a, long b) {
    long r;
    if (__builtin_add_overflow(a, b, &r)) return a > 0 ? LONG_MAX : LONG_MIN;
    return r;
}

// atomics of a serial run, and bit reinterpretation
template <class T, class U> inline T atomic_or(T *p, U v) { T old = *p; *p = old | (T) v; return old; }
template <class T, class U> inline T atomic_max(T *p, U v) { T old = *p; if ((T) v > old) *p = (T) v; return old; }
template <class T, class U> inline T atomic_add(T *p, U v) { T old = *p; *p = old + (T) v; return old; }
inline int as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

// a growing output list of the list-of-lists builder
template <class T> struct plb_list {
    std::vector<T> items;
    bool omitted = false;
    inline void append(T v) {
        if (!omitted) items.push_back(v);
    }
};
