// In-register FFT codelets for the multi-view deconvolution passes (sm_100a).
//
// Everything on this path is a complex FFT of a length N = 2^a 3^b 5^c that is split into 2 or 3
// "stages"; in each stage one thread runs a radix-R DFT (R <= 36) entirely in registers.  The
// transforms are *no-reorder*: the forward transform is decimation-in-frequency and leaves the
// spectrum in a fixed, scrambled (digit-reversed) order; the inverse is the exact conjugate
// transpose of the forward flow graph and consumes that order.  Because every operation between a
// forward and its inverse on this path is point-wise in frequency (spectrum multiply with a kernel
// spectrum that was produced by the very same forward transform), the scrambled order never has to
// be undone -- this removes one shared-memory pass per FFT.
//
// The header is plain C++17 and also compiles with g++ (tests/host emulation of the kernels).
#pragma once
#include <type_traits>
#include <utility>

#if defined(__CUDACC__)
#define MVD_HD __host__ __device__ __forceinline__
#else
#define MVD_HD inline __attribute__((always_inline))
#endif

namespace mvd {

struct alignas(8) cpx { float x, y; };

// Complex arithmetic.  On the device a complex number travels as ONE 64-bit register pair (re, im) through the packed f32x2
// instructions of sm_100a (FADD2 / FMUL2 / FFMA2: two IEEE single-precision lanes per issue slot).  ptxas folds the half swaps,
// per-half negations and scalar broadcasts written below as mov.b64 packing into operand modifiers (.LO_HI, .NP, .F32), so
//   a + b, a - b, a -+ i b       : 1 instruction   (2 scalar)
//   c * a, c * a + b  (c real)   : 1               (2)
//   a * b, a * conj(b)           : 2               (4)
// The host build (emulator, tests) uses the plain scalar formulas.
#if defined(__CUDA_ARCH__)
typedef unsigned long long pk64;
MVD_HD pk64 pk2(float lo, float hi) { pk64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
MVD_HD pk64 pkc(cpx a) { return pk2(a.x, a.y); }
MVD_HD cpx upk(pk64 v) { cpx r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
MVD_HD pk64 p_add(pk64 a, pk64 b) { pk64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
MVD_HD pk64 p_mul(pk64 a, pk64 b) { pk64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
MVD_HD pk64 p_fma(pk64 a, pk64 b, pk64 c) { pk64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
MVD_HD cpx operator+(cpx a, cpx b) { return upk(p_add(pkc(a), pkc(b))); }
MVD_HD cpx operator-(cpx a, cpx b) { return upk(p_add(pkc(a), pk2(-b.x, -b.y))); }
MVD_HD cpx c_add_mi(cpx a, cpx b) { return upk(p_add(pkc(a), pk2(b.y, -b.x))); }          // a - i b
MVD_HD cpx c_add_pi(cpx a, cpx b) { return upk(p_add(pkc(a), pk2(-b.y, b.x))); }          // a + i b
MVD_HD cpx c_scale(float c, cpx a) { return upk(p_mul(pkc(a), pk2(c, c))); }              // c a
MVD_HD cpx c_fma(float c, cpx a, cpx b) { return upk(p_fma(pkc(a), pk2(c, c), pkc(b))); } // c a + b
MVD_HD cpx c_fma_mi(float c, cpx a, cpx b) { return upk(p_fma(pk2(a.y, a.x), pk2(c, -c), pkc(b))); }   // b - i c a
MVD_HD cpx c_fma_pi(float c, cpx a, cpx b) { return upk(p_fma(pk2(a.y, a.x), pk2(-c, c), pkc(b))); }   // b + i c a
MVD_HD cpx cmul(cpx a, cpx b) { return upk(p_fma(pk2(a.y, a.x), pk2(-b.y, b.y), p_mul(pkc(a), pk2(b.x, b.x)))); }
MVD_HD cpx cmul_conj(cpx a, cpx b) { return upk(p_fma(pk2(a.y, a.x), pk2(b.y, -b.y), p_mul(pkc(a), pk2(b.x, b.x)))); }  // a * conj(b)
#else
MVD_HD cpx operator+(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
MVD_HD cpx operator-(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
MVD_HD cpx c_add_mi(cpx a, cpx b) { return cpx{a.x + b.y, a.y - b.x}; }
MVD_HD cpx c_add_pi(cpx a, cpx b) { return cpx{a.x - b.y, a.y + b.x}; }
MVD_HD cpx c_scale(float c, cpx a) { return cpx{c * a.x, c * a.y}; }
MVD_HD cpx c_fma(float c, cpx a, cpx b) { return cpx{c * a.x + b.x, c * a.y + b.y}; }
MVD_HD cpx c_fma_mi(float c, cpx a, cpx b) { return cpx{b.x + c * a.y, b.y - c * a.x}; }
MVD_HD cpx c_fma_pi(float c, cpx a, cpx b) { return cpx{b.x - c * a.y, b.y + c * a.x}; }
MVD_HD cpx cmul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
MVD_HD cpx cmul_conj(cpx a, cpx b) { return cpx{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }  // a * conj(b)
#endif

// L2 prefetch of the 128-byte line containing p (no-op on the host)
#if defined(__CUDA_ARCH__)
MVD_HD void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
MVD_HD void prefetch_l2(const void*) {}
#endif

// read-only / streaming loads
#if defined(__CUDA_ARCH__)
MVD_HD cpx ld_ro(const cpx* p) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return cpx{v.x, v.y}; }
MVD_HD float ld_rof(const float* p) { return __ldg(p); }
MVD_HD cpx ld_stream(const cpx* p) { const float2 v = *reinterpret_cast<const float2*>(p); return cpx{v.x, v.y}; }
#else
MVD_HD cpx ld_ro(const cpx* p) { return *p; }
MVD_HD float ld_rof(const float* p) { return *p; }
MVD_HD cpx ld_stream(const cpx* p) { return *p; }
#endif

// ---------------------------------------------------------------------------------------------
// compile-time trigonometry (double precision, exact at the octant points)
// ---------------------------------------------------------------------------------------------
constexpr double kTwoPi = 6.283185307179586476925286766559005768;

constexpr double cx_sin_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i <= 12; ++i) { term *= -x2 / double((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double cx_cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i <= 12; ++i) { term *= -x2 / double((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
struct cx_pair { double c, s; };
// cos/sin of 2*pi*p/q for 0 <= p/q <= 1
constexpr cx_pair cx_cossin_turn(long long p, long long q) {
    p %= q; if (p < 0) p += q;
    if (2 * p > q) { cx_pair r = cx_cossin_turn(q - p, q); return cx_pair{r.c, -r.s}; }          // > half turn
    if (4 * p > q) { cx_pair r = cx_cossin_turn(q - 2 * p, 2 * q); return cx_pair{-r.c, r.s}; }   // (1/4, 1/2]
    if (8 * p > q) { cx_pair r = cx_cossin_turn(q - 4 * p, 4 * q); return cx_pair{r.s, r.c}; }    // (1/8, 1/4]
    double x = kTwoPi * double(p) / double(q);
    return cx_pair{cx_cos_small(x), cx_sin_small(x)};
}

constexpr int cx_gcd(int a, int b) { return b == 0 ? a : cx_gcd(b, a % b); }

// multiply by the compile-time twiddle exp(-+ 2 pi i K / N)  (INV -> conjugate)
template <int K, int N, bool INV>
MVD_HD cpx mul_tw(cpx a) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (2 * k == N) {
        return cpx{-a.x, -a.y};
    } else if constexpr (4 * k == N) {          // forward: * (-i)
        return INV ? cpx{-a.y, a.x} : cpx{a.y, -a.x};
    } else if constexpr (4 * k == 3 * N) {      // forward: * (+i)
        return INV ? cpx{a.y, -a.x} : cpx{-a.y, a.x};
    } else if constexpr ((8 * k) % N == 0) {    // odd multiples of 1/8 turn
        constexpr float h = 0.70710678118654752440f;
        constexpr int o = (8 * k) / N;          // 1,3,5,7
        // forward twiddle = cos - i sin
        constexpr float cs = (o == 1 || o == 7) ? h : -h;
        constexpr float sn0 = (o == 1 || o == 3) ? -h : h;       // imaginary part (forward)
        constexpr float sn = INV ? -sn0 : sn0;
        return c_fma_pi(sn, a, c_scale(cs, a));                  // (cs + i sn) a
    } else {
        constexpr cx_pair t = cx_cossin_turn(k, N);
        constexpr float c = float(t.c);
        constexpr float s = float(INV ? t.s : -t.s);
        return c_fma_pi(s, a, c_scale(c, a));                    // (c + i s) a
    }
}

template <int I, int N, class F>
MVD_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// ---------------------------------------------------------------------------------------------
// prime butterflies (natural order in -> natural order out, in place)
// ---------------------------------------------------------------------------------------------
template <bool INV> MVD_HD void bfly2(cpx& a, cpx& b) {
    cpx t = a - b; a = a + b; b = t;
}
template <bool INV> MVD_HD void bfly3(cpx& a, cpx& b, cpx& c) {
    constexpr float s = 0.86602540378443864676f;
    const cpx t = b + c;
    const cpx d = b - c;
    const cpx u = c_fma(-0.5f, t, a);
    a = a + t;
    // forward: X1 = u - i s d ; X2 = u + i s d
    if constexpr (!INV) { b = c_fma_mi(s, d, u); c = c_fma_pi(s, d, u); }
    else { b = c_fma_pi(s, d, u); c = c_fma_mi(s, d, u); }
}
template <bool INV> MVD_HD void bfly4(cpx& a, cpx& b, cpx& c, cpx& d) {
    const cpx t0 = a + c, t1 = a - c, t2 = b + d, t3 = b - d;
    a = t0 + t2; c = t0 - t2;
    if constexpr (!INV) { b = c_add_mi(t1, t3); d = c_add_pi(t1, t3); }     // forward: t1 -+ i t3
    else { b = c_add_pi(t1, t3); d = c_add_mi(t1, t3); }
}
template <bool INV> MVD_HD void bfly5(cpx& a, cpx& b, cpx& c, cpx& d, cpx& e) {
    constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const cpx t1 = b + e, t2 = c + d, t3 = b - e, t4 = c - d;
    const cpx m1 = c_fma(c2, t2, c_fma(c1, t1, a));
    const cpx m2 = c_fma(c1, t2, c_fma(c2, t1, a));
    const cpx n1 = c_fma(s2, t4, c_scale(s1, t3));
    const cpx n2 = c_fma(-s1, t4, c_scale(s2, t3));
    a = a + t1 + t2;
    // forward: X1 = m1 - i n1, X4 = m1 + i n1, X2 = m2 - i n2, X3 = m2 + i n2
    if constexpr (!INV) { b = c_add_mi(m1, n1); e = c_add_pi(m1, n1); c = c_add_mi(m2, n2); d = c_add_pi(m2, n2); }
    else { b = c_add_pi(m1, n1); e = c_add_mi(m1, n1); c = c_add_pi(m2, n2); d = c_add_mi(m2, n2); }
}

constexpr int first_factor(int R) {
    return R % 4 == 0 ? 4 : R % 2 == 0 ? 2 : R % 3 == 0 ? 3 : R % 5 == 0 ? 5 : R;
}
constexpr bool radix_supported(int R) {
    while (R % 2 == 0) R /= 2;
    while (R % 3 == 0) R /= 3;
    while (R % 5 == 0) R /= 5;
    return R == 1;
}
// Coprime split R = r * m for the prime-factor (Good-Thomas) step: r = the whole power of the smallest prime of R, or 0 when R is a
// prime power.  With n = (m n1 + r n2) mod R the length-R DFT is an r x m two-dimensional DFT WITHOUT twiddle factors between the
// two steps; the output frequency k is given by k = k1 mod r, k = k2 mod m (Chinese remainder theorem).
constexpr int pfa_factor(int R) {
    int p = R % 2 == 0 ? 2 : R % 3 == 0 ? 3 : R % 5 == 0 ? 5 : R;
    int r = 1;
    while (R % p == 0) { r *= p; R /= p; }
    return R == 1 ? 0 : r;
}
// frequency index held at position p after the no-reorder forward DFT of length R
constexpr int freq_of_pos(int R, int p) {
    if (R == 1) return 0;
    const int rp = pfa_factor(R);
    if (rp == 0) {                                   // prime power: decimation in frequency with twiddles
        int r = first_factor(R), m = R / r;
        return (p / m) + r * freq_of_pos(m, p % m);
    }
    const int r = rp, m = R / rp;
    for (int t = 0; t < r; ++t)
        for (int s = 0; s < m; ++s)
            if ((m * t + r * s) % R == p) {
                const int k1 = freq_of_pos(r, t), k2 = freq_of_pos(m, s);
                for (int k = 0; k < R; ++k)
                    if (k % r == k1 && k % m == k2) return k;
            }
    return -1;
}

// ---------------------------------------------------------------------------------------------
// recursive in-register DFT over the elements a[Map::at(i)], i in [0,R)
//   prime powers (Cooley-Tukey, decimation in frequency):
//     forward: (I_r (x) F_m) . D . (B_r (x) I_m)      (output scrambled, see freq_of_pos)
//   coprime factors (prime-factor algorithm): F_r along n1 then F_m along n2 of n = (m n1 + r n2) mod R, no twiddles --
//     radix 6, 10, 12, 15, 18, 20, 24, 30 save all their internal twiddle multiplications
//   inverse: exact conjugate transpose of the forward flow graph (unnormalised)
// ---------------------------------------------------------------------------------------------
template <int OFF, int STR> struct MapLin { static constexpr int at(int i) { return OFF + i * STR; } };
template <class M, int Q, int LEN> struct MapBlock { static constexpr int at(int i) { return M::at(Q * LEN + i); } };           // Cooley-Tukey sub-block
template <class M, int R, int r, int m, int N2> struct MapPfa1 { static constexpr int at(int i) { return M::at((m * i + r * N2) % R); } };   // along n1
template <class M, int R, int r, int m, int T> struct MapPfa2 { static constexpr int at(int i) { return M::at((m * T + r * i) % R); } };     // along n2

template <int R, class Map, bool INV, int E>
struct DftM {
    static MVD_HD void run(cpx (&a)[E]) {
        if constexpr (R > 1) {
            static_assert(radix_supported(R), "radix must be 2^a 3^b 5^c");
            constexpr int rp = pfa_factor(R);
            if constexpr (rp != 0) {
                constexpr int r = rp, m = R / rp;
                if constexpr (!INV) {
                    static_for<0, m>([&](auto c) { DftM<r, MapPfa1<Map, R, r, m, decltype(c)::value>, false, E>::run(a); });
                    static_for<0, r>([&](auto c) { DftM<m, MapPfa2<Map, R, r, m, decltype(c)::value>, false, E>::run(a); });
                } else {
                    static_for<0, r>([&](auto c) { DftM<m, MapPfa2<Map, R, r, m, decltype(c)::value>, true, E>::run(a); });
                    static_for<0, m>([&](auto c) { DftM<r, MapPfa1<Map, R, r, m, decltype(c)::value>, true, E>::run(a); });
                }
            } else {
                constexpr int r = first_factor(R);
                constexpr int m = R / r;
                if constexpr (!INV) {
                    static_for<0, m>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        bfly<r, j, m>(a);
                        static_for<1, r>([&](auto qc) {
                            constexpr int q = decltype(qc)::value;
                            a[Map::at(j + q * m)] = mul_tw<j * q, R, false>(a[Map::at(j + q * m)]);
                        });
                    });
                    static_for<0, r>([&](auto qc) { DftM<m, MapBlock<Map, decltype(qc)::value, m>, false, E>::run(a); });
                } else {
                    static_for<0, r>([&](auto qc) { DftM<m, MapBlock<Map, decltype(qc)::value, m>, true, E>::run(a); });
                    static_for<0, m>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        static_for<1, r>([&](auto qc) {
                            constexpr int q = decltype(qc)::value;
                            a[Map::at(j + q * m)] = mul_tw<j * q, R, true>(a[Map::at(j + q * m)]);
                        });
                        bfly<r, j, m>(a);
                    });
                }
            }
        }
    }
    template <int r, int j, int m>
    static MVD_HD void bfly(cpx (&a)[E]) {
        if constexpr (r == 2) bfly2<INV>(a[Map::at(j)], a[Map::at(j + m)]);
        else if constexpr (r == 3) bfly3<INV>(a[Map::at(j)], a[Map::at(j + m)], a[Map::at(j + 2 * m)]);
        else if constexpr (r == 4) bfly4<INV>(a[Map::at(j)], a[Map::at(j + m)], a[Map::at(j + 2 * m)], a[Map::at(j + 3 * m)]);
        else bfly5<INV>(a[Map::at(j)], a[Map::at(j + m)], a[Map::at(j + 2 * m)], a[Map::at(j + 3 * m)], a[Map::at(j + 4 * m)]);
    }
};
template <int R, int OFF, int STR, bool INV, int E>
struct Dft {
    static MVD_HD void run(cpx (&a)[E]) { DftM<R, MapLin<OFF, STR>, INV, E>::run(a); }
};

// R consecutive complex values <-> global memory; 16-byte vector accesses when R is even (callers guarantee 16-byte
// alignment in that case: line pitch is a multiple of 4 elements and the element offset a multiple of R)
template <int R>
MVD_HD void st_vec(cpx* o, const cpx (&a)[R]) {
#if defined(__CUDA_ARCH__)
    if constexpr (R % 2 == 0) {
        static_for<0, R / 2>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            reinterpret_cast<float4*>(o)[i] = make_float4(a[2 * i].x, a[2 * i].y, a[2 * i + 1].x, a[2 * i + 1].y);
        });
        return;
    }
#endif
    static_for<0, R>([&](auto ic) { constexpr int i = decltype(ic)::value; o[i] = a[i]; });
}
template <int R>
MVD_HD void ld_vec(const cpx* o, cpx (&a)[R]) {
#if defined(__CUDA_ARCH__)
    if constexpr (R % 2 == 0) {
        static_for<0, R / 2>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const float4 v = reinterpret_cast<const float4*>(o)[i];
            a[2 * i] = cpx{v.x, v.y}; a[2 * i + 1] = cpx{v.z, v.w};
        });
        return;
    }
#endif
    static_for<0, R>([&](auto ic) { constexpr int i = decltype(ic)::value; a[i] = o[i]; });
}

// ---------------------------------------------------------------------------------------------
// FFT plan: N = R1*R2*R3 (R3 == 1 for two-stage plans).
//   column passes (y/z axis): T threads cooperate on one column, W columns (128-byte row segments) per CTA
//   x passes (contiguous lines): XT threads cooperate on one line, XL lines per CTA
// ---------------------------------------------------------------------------------------------
template <int N_, int R1_, int R2_, int R3_, int T_, int W_, int XT_ = N_ / R1_, int XL_ = 8>
struct Plan {
    static constexpr int N = N_, R1 = R1_, R2 = R2_, R3 = R3_, T = T_, W = W_, XT = XT_, XL = XL_;
    static constexpr int NSTAGES = (R3_ > 1) ? 3 : 2;
    static constexpr int BLK1 = N_, BLK2 = N_ / R1_, BLK3 = N_ / (R1_ * R2_);
    static constexpr int THREADS = T_ * W_;
    static constexpr int XTHREADS = XT_ * XL_;
    static_assert(R1_ * R2_ * R3_ == N_, "plan radices must multiply to N");
    static_assert(R2_ > 1, "plans have at least two stages");
    static_assert(radix_supported(R1_) && radix_supported(R2_) && radix_supported(R3_), "unsupported radix");
};

// multiply a[1..R) by per-position twiddles twp(p) (conjugated for the inverse)
template <int R, bool INV, class TwP>
MVD_HD void apply_tw(cpx (&a)[R], TwP&& twp) {
    static_for<1, R>([&](auto pc) {
        constexpr int p = decltype(pc)::value;
        a[p] = INV ? cmul_conj(a[p], twp(pc)) : cmul(a[p], twp(pc));   // twp receives an integral_constant
    });
}

// One radix-R butterfly of one stage.  The line is accessed through functors:
//   src(base, off) -> cpx, dst(base, off, cpx) for sample base + off (off a compile-time constant);
//   tw(pc, j) -> the stage twiddle of butterfly position p at index j, exp(-2 pi i (N/BLK) freq_of_pos(R,p) j / N), from a
//   per-position table [p-1][j] (global or shared memory): one base address per butterfly, compile-time offsets per position.
// BLK = block length handled by this stage (N for the first stage), S = BLK/R the element stride.
// Position p of the butterfly (element base + p*S) holds frequency freq_of_pos(R,p) of the radix-R DFT.
template <int N, int BLK, int R, bool INV, class TwF, class Src, class Dst>
MVD_HD void stage_bfly(int g, TwF&& tw, Src&& src, Dst&& dst) {
    constexpr int S = BLK / R;
    const int b = g / S;
    const int j = g - b * S;
    const int base = b * BLK + j;
    cpx a[R];
    static_for<0, R>([&](auto pc) { constexpr int p = decltype(pc)::value; a[p] = src(base, p * S); });
    auto twp = [&](auto pc) { return tw(pc, j); };
    if constexpr (!INV) {
        Dft<R, 0, 1, false, R>::run(a);
        if constexpr (S > 1) apply_tw<R, false>(a, twp);
    } else {
        if constexpr (S > 1) apply_tw<R, true>(a, twp);
        Dft<R, 0, 1, true, R>::run(a);
    }
    static_for<0, R>([&](auto pc) { constexpr int p = decltype(pc)::value; dst(base, p * S, a[p]); });
}

// last forward stage (S == 1) fused with the spectrum multiply and the first inverse stage:
//   a <- F_R a ;  a[p] *= khat(base + p) ;  a <- F_R^H a
template <int N, int BLK, int R, class Src, class Dst, class Khat>
MVD_HD void stage_conv(int g, Src&& src, Dst&& dst, Khat&& khat) {
    static_assert(BLK == R, "stage_conv is the last forward stage");
    const int base = g * BLK;
    cpx a[R], kh[R];
    // the kernel-spectrum loads go out first: their (L2 / DRAM) latency overlaps the shared-memory reads and the forward DFT
    static_for<0, R>([&](auto pc) { constexpr int p = decltype(pc)::value; kh[p] = khat(base, p); });
    static_for<0, R>([&](auto pc) { constexpr int p = decltype(pc)::value; a[p] = src(base, p); });
    Dft<R, 0, 1, false, R>::run(a);
    static_for<0, R>([&](auto pc) { constexpr int p = decltype(pc)::value; a[p] = cmul(a[p], kh[p]); });
    Dft<R, 0, 1, true, R>::run(a);
    static_for<0, R>([&](auto pc) { constexpr int p = decltype(pc)::value; dst(base, p, a[p]); });
}

// frequency index (natural DFT order) stored at line position n after the full forward plan.
template <class P>
constexpr int plan_freq_of_pos(int n) {
    // stage 1: n = j1 + p1*S1 ... generic: position digits (p1, p2, p3) with
    // n = p1*BLK2 + p2*BLK3 + p3 ; frequency = f1 + R1*(f2 + R2*f3)
    int p1 = n / P::BLK2, rem = n % P::BLK2;
    int p2 = rem / P::BLK3, p3 = rem % P::BLK3;
    int f1 = freq_of_pos(P::R1, p1), f2 = freq_of_pos(P::R2, p2);
    int f3 = (P::R3 > 1) ? freq_of_pos(P::R3, p3) : 0;
    return f1 + P::R1 * (f2 + P::R2 * f3);
}

}  // namespace mvd
