// In-register DFT codelets, FP64, natural-order in/out, fully unrolled at compile time.
//
// dft<R, INV>(v):  v[p] <- sum_k v[k] * w^(k p),  w = exp(-2 pi i / R)  (INV = false, "forward",
// CPMD isign=+1, mltfft_utils.mod.F90:514-518) or exp(+2 pi i / R) (INV = true, "inverse",
// isign=-1).  No scaling.
//
// Supported radices: 2,3,4,5,7 (direct butterflies) and the composites listed in Split<>
// (Cooley-Tukey in registers with literal twiddles from roots.h; trivial twiddles cost nothing).
// This plays the role of the reference's radix-3/4/5/6/8 passes (gfft_utils.mod.F90:136-3412)
// but is a different algorithm: the reference runs Stockham passes over a cache-blocked batch in
// memory; here one thread owns a whole radix-R sub-transform in registers.
#pragma once
#include "cpb_defs.h"
#include "roots.h"

namespace cpb {

// multiply by exp(-/+ 2 pi i K / DEN)
template <int DEN, int K, bool INV>
CPB_HD cplx mul_root(cplx a) {
  constexpr int k = ((K % DEN) + DEN) % DEN;
  if constexpr (k == 0) {
    return a;
  } else if constexpr (2 * k == DEN) {
    return mk(-a.x, -a.y);
  } else if constexpr (4 * k == DEN) {
    // forward: * (-i) ; inverse: * (+i)
    return INV ? mk(-a.y, a.x) : mk(a.y, -a.x);
  } else if constexpr (4 * k == 3 * DEN) {
    return INV ? mk(a.y, -a.x) : mk(-a.y, a.x);
  } else {
    constexpr double c = Root<DEN>::c(k);
    constexpr double s = INV ? Root<DEN>::s(k) : -Root<DEN>::s(k);
    return mk(a.x * c - a.y * s, a.x * s + a.y * c);
  }
}

template <int R>
struct Split {
  static constexpr int a = 1;  // 1 => prime / direct butterfly
};
#define CPB_SPLIT(R, A) \
  template <>           \
  struct Split<R> {     \
    static constexpr int a = A; \
  };
CPB_SPLIT(6, 2)
CPB_SPLIT(8, 2)
CPB_SPLIT(9, 3)
CPB_SPLIT(10, 2)
CPB_SPLIT(12, 4)
CPB_SPLIT(14, 2)
CPB_SPLIT(15, 3)
CPB_SPLIT(16, 4)
CPB_SPLIT(18, 2)
CPB_SPLIT(20, 4)
CPB_SPLIT(21, 3)
CPB_SPLIT(24, 4)
CPB_SPLIT(25, 5)
CPB_SPLIT(28, 4)
CPB_SPLIT(30, 5)
CPB_SPLIT(32, 4)
#undef CPB_SPLIT

template <int R, bool INV>
CPB_HD void dft(cplx (&v)[R]);

// odd prime P: pair up k and P-k
template <int P, bool INV>
CPB_HD void dft_prime(cplx (&v)[P]) {
  constexpr int H = (P - 1) / 2;
  cplx s[H + 1], d[H + 1];
  static_for<1, H + 1>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    s[j] = cadd(v[j], v[P - j]);
    d[j] = csub(v[j], v[P - j]);
  });
  cplx v0 = v[0];
  cplx sum = v0;
  static_for<1, H + 1>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    sum = cadd(sum, s[j]);
  });
  v[0] = sum;
  static_for<1, H + 1>([&](auto kk) {
    constexpr int k = decltype(kk)::value;
    cplx A = v0;
    cplx Bv = mk(0.0, 0.0);
    static_for<1, H + 1>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      constexpr double c = Root<P>::c((j * k) % P);
      constexpr double sn = Root<P>::s((j * k) % P);
      A.x += c * s[j].x;
      A.y += c * s[j].y;
      if constexpr (j == 1) {
        Bv.x = sn * d[j].x;
        Bv.y = sn * d[j].y;
      } else {
        Bv.x += sn * d[j].x;
        Bv.y += sn * d[j].y;
      }
    });
    // forward: X_k = A - i B, X_{P-k} = A + i B ; inverse: swapped
    cplx mib = mk(Bv.y, -Bv.x);  // -i * B
    if (INV) {
      v[k] = csub(A, mib);
      v[P - k] = cadd(A, mib);
    } else {
      v[k] = cadd(A, mib);
      v[P - k] = csub(A, mib);
    }
  });
}

template <bool INV>
CPB_HD void dft4(cplx (&v)[4]) {
  cplx t0 = cadd(v[0], v[2]);
  cplx t1 = csub(v[0], v[2]);
  cplx t2 = cadd(v[1], v[3]);
  cplx t3 = mul_root<4, 1, INV>(csub(v[1], v[3]));
  v[0] = cadd(t0, t2);
  v[1] = cadd(t1, t3);
  v[2] = csub(t0, t2);
  v[3] = csub(t1, t3);
}

// Good-Thomas (prime factor) split R = Ra * Rb with coprime factors: input index (Rb n1 + Ra n2) mod R, output
// index k with k mod Ra = k1 and k mod Rb = k2 - a plain 2-D transform, no twiddle factors between the stages
// (the Cooley-Tukey split of 12 = 4 x 3 spends 16 of its 120 FP64 instructions on them).
template <int R>
struct Pfa {
  static constexpr int a = 0;  // 0 => no prime-factor split
};
#define CPB_PFA(R, A)          \
  template <>                  \
  struct Pfa<R> {              \
    static constexpr int a = A; \
  };
#ifndef CPB_NO_PFA
CPB_PFA(6, 2)
CPB_PFA(10, 2)
CPB_PFA(12, 4)
CPB_PFA(14, 2)
CPB_PFA(15, 3)
CPB_PFA(18, 2)
CPB_PFA(20, 4)
CPB_PFA(21, 3)
CPB_PFA(24, 3)
CPB_PFA(28, 4)
CPB_PFA(30, 5)
#endif
#undef CPB_PFA
CPB_HD constexpr int cpb_crt(int k1, int ra, int k2, int rb) {
  for (int k = 0; k < ra * rb; ++k)
    if (k % ra == k1 && k % rb == k2) return k;
  return -1;
}

template <int R, bool INV>
CPB_HD void dft(cplx (&v)[R]) {
  if constexpr (Pfa<R>::a != 0) {
    constexpr int Ra = Pfa<R>::a;
    constexpr int Rb = R / Ra;
    cplx t[R];
    static_for<0, Rb>([&](auto nn) {
      constexpr int n2 = decltype(nn)::value;
      cplx u[Ra];
      static_for<0, Ra>([&](auto mm) {
        constexpr int n1 = decltype(mm)::value;
        u[n1] = v[(Rb * n1 + Ra * n2) % R];
      });
      dft<Ra, INV>(u);
      static_for<0, Ra>([&](auto kk) {
        constexpr int k1 = decltype(kk)::value;
        t[k1 * Rb + n2] = u[k1];
      });
    });
    static_for<0, Ra>([&](auto kk) {
      constexpr int k1 = decltype(kk)::value;
      cplx w[Rb];
      static_for<0, Rb>([&](auto nn) {
        constexpr int n2 = decltype(nn)::value;
        w[n2] = t[k1 * Rb + n2];
      });
      dft<Rb, INV>(w);
      static_for<0, Rb>([&](auto qq) {
        constexpr int k2 = decltype(qq)::value;
        v[cpb_crt(k1, Ra, k2, Rb)] = w[k2];
      });
    });
  } else if constexpr (R == 1) {
  } else if constexpr (R == 2) {
    cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if constexpr (R == 4) {
    dft4<INV>(v);
  } else if constexpr (Split<R>::a == 1) {
    static_assert(R == 3 || R == 5 || R == 7, "unsupported prime radix");
    dft_prime<R, INV>(v);
  } else {
    // Cooley-Tukey R = Ra * Rb: index k = j + Rb*ka, output p + Ra*q
    constexpr int Ra = Split<R>::a;
    constexpr int Rb = R / Ra;
    cplx t[R];
    static_for<0, Rb>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      cplx u[Ra];
      static_for<0, Ra>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        u[k] = v[j + Rb * k];
      });
      dft<Ra, INV>(u);
      static_for<0, Ra>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        t[j * Ra + p] = mul_root<R, j * p, INV>(u[p]);
      });
    });
    static_for<0, Ra>([&](auto pp) {
      constexpr int p = decltype(pp)::value;
      cplx u[Rb];
      static_for<0, Rb>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        u[j] = t[j * Ra + p];
      });
      dft<Rb, INV>(u);
      static_for<0, Rb>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        v[p + Ra * q] = u[q];
      });
    });
  }
}

// ---------------------------------------------------------------------------------------------
// dft_in<R, INV, LO, HI>(v): same transform, but the caller guarantees v[k] == 0 for k outside
// [LO, HI) (those entries are ignored, not read).  Every 1-D transform of the sparse wavefunction
// FFT has this structure: the G-sphere only fills the middle half of each axis
// (fftprp_utils.mod.F90:161-192), so the first radix pass sees zeros in half of its inputs.
// Zero terms are dropped at compile time (the compiler may not fold x + 0.0 under IEEE rules).
// ---------------------------------------------------------------------------------------------
template <int R>
struct LeafDirectMax {  // largest non-zero count for which the direct sum beats the butterfly
  static constexpr int v = (R == 4) ? 2 : 1;
};

CPB_HD constexpr int cpb_ceil_div(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }
CPB_HD constexpr int cpb_clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

template <int R, bool INV, int LO, int HI>
CPB_HD void dft_in(cplx (&v)[R]) {
  if constexpr (LO <= 0 && HI >= R) {
    dft<R, INV>(v);
  } else if constexpr (HI <= LO) {
    static_for<0, R>([&](auto pp) { v[decltype(pp)::value] = mk(0.0, 0.0); });
  } else if constexpr (Split<R>::a == 1 || R == 4 || R == 2) {
    if constexpr (HI - LO <= LeafDirectMax<R>::v) {
      cplx o[R];
      static_for<0, R>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        o[p] = mul_root<R, LO * p, INV>(v[LO]);
        static_for<LO + 1, HI>([&](auto kk) {
          constexpr int k = decltype(kk)::value;
          o[p] = cadd(o[p], mul_root<R, k * p, INV>(v[k]));
        });
      });
      static_for<0, R>([&](auto pp) { v[decltype(pp)::value] = o[decltype(pp)::value]; });
    } else {
      static_for<0, R>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k < LO || k >= HI) v[k] = mk(0.0, 0.0);
      });
      dft<R, INV>(v);
    }
  } else {
    constexpr int Ra = Split<R>::a;
    constexpr int Rb = R / Ra;
    cplx t[R];
    static_for<0, Rb>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      // non-zero ka: LO <= j + Rb*ka < HI
      constexpr int ka_lo = cpb_clamp(cpb_ceil_div(LO - j, Rb), 0, Ra);
      constexpr int ka_hi = cpb_clamp(cpb_ceil_div(HI - j, Rb), 0, Ra);
      cplx u[Ra];
      static_for<0, Ra>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        u[k] = v[j + Rb * k];
      });
      dft_in<Ra, INV, ka_lo, ka_hi>(u);
      static_for<0, Ra>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        t[j * Ra + p] = mul_root<R, j * p, INV>(u[p]);
      });
    });
    static_for<0, Ra>([&](auto pp) {
      constexpr int p = decltype(pp)::value;
      cplx u[Rb];
      static_for<0, Rb>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        u[j] = t[j * Ra + p];
      });
      dft<Rb, INV>(u);
      static_for<0, Rb>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        v[p + Ra * q] = u[q];
      });
    });
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming forms of the band-pruned transforms, R = D * M (used by the warp-autonomous z kernels,
// kernels_zw.h).  They touch the R-point transform in D groups of M, so that only the band elements
// and one M-point working set are live in registers at a time.
//
// dft_in_dif: decimation in frequency over the output index p = D*m + r.  v[k] is non-zero only for
// k in [LO, HI).   g_r[u] = sum_c v[u + M c] w_R^((u + M c) r),   Y[D m + r] = DFT_M(g_r)[m];
// emit(IC<p>, Y[p]) is called group by group (r = 0 .. D-1), so the caller can twiddle and store a
// group before the next one is computed.
// ---------------------------------------------------------------------------------------------
template <int R, int D, bool INV, int LO, int HI, class E>
CPB_HD void dft_in_dif(const cplx (&v)[R], E&& emit) {
  static_assert(R % D == 0, "R = D * M");
  constexpr int M = R / D;
  static_for<0, D>([&](auto rr) {
    constexpr int r = decltype(rr)::value;
    cplx g[M];
    static_for<0, M>([&](auto uu) {
      constexpr int u = decltype(uu)::value;
      bool first = true;  // folded at compile time (every branch below is constexpr)
      g[u] = mk(0.0, 0.0);
      static_for<0, D>([&](auto cc) {
        constexpr int k = u + M * decltype(cc)::value;
        if constexpr (k >= LO && k < HI) {
          const cplx t = mul_root<R, k * r, INV>(v[k]);
          g[u] = first ? t : cadd(g[u], t);
          first = false;
        }
      });
    });
    dft<M, INV>(g);
    static_for<0, M>([&](auto mm) {
      constexpr int m = decltype(mm)::value;
      emit(IC<D * m + r>{}, g[m]);
    });
  });
}

// dft_out_dit: decimation in time over the input index p = D*m + r, only the outputs t in [LO, HI)
// are produced (the others of z are left untouched).  in(IC<p>) returns input p.
//   E_r = DFT_M(in(D m + r)),   Z[t] = sum_r w_R^(r t) E_r[t mod M]
template <int R, int D, bool INV, int LO, int HI, class IN>
CPB_HD void dft_out_dit(IN&& in, cplx (&z)[R]) {
  static_assert(R % D == 0, "R = D * M");
  constexpr int M = R / D;
  static_for<0, D>([&](auto rr) {
    constexpr int r = decltype(rr)::value;
    cplx e[M];
    static_for<0, M>([&](auto mm) {
      constexpr int m = decltype(mm)::value;
      e[m] = in(IC<D * m + r>{});
    });
    dft<M, INV>(e);
    static_for<LO, HI>([&](auto tt) {
      constexpr int t = decltype(tt)::value;
      const cplx c = mul_root<R, r * t, INV>(e[t % M]);
      if constexpr (r == 0) z[t] = c;
      else z[t] = cadd(z[t], c);
    });
  });
}

}  // namespace cpb
