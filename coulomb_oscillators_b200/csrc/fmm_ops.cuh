// fmm_ops.cuh -- 3D Cartesian FMM operator algebra, compile-time unrolled in the order P.
//
// Maths restated from SURVEY.md section 2.4 (reference Simulation/fmm_cart_base3.cuh):
//   storage      symmetric order-n tensor A[x,y,z], x+y+z = n, at i(x,z) = (n(n+1)-(n-z)(n-z+1))/2 + n - x
//                (:210-213); tuple of orders 0..q-1 starts order q at q(q+1)(q+2)/6 (:180-183);
//                traceless tensors keep only z in {0,1}: i(x,z) = (z+1)n - x, order q at q^2 (:185-232)
//   P2M          M_q[x,y,z] += (-1)^q/q! dx^x dy^y dz^z                                  (:951-961)
//   M2M          M'_n[x,y,z] += 1/n! sum_m (n-m)! sum_k C(x,k1)C(y,k2)C(z,k3) d^k M_{n-m}[x-k1,y-k2,z-k3] (:1111-1146)
//   M2L          L_n += 1/n! <M_k, grad^{n+k}(1/r)>, n >= 1, k != 1, n+k <= P            (:1265-1346)
//   grad^m(1/r)  closed form for z in {0,1} (:768-804), z >= 2 from the trace relation   (:644-659)
//   L2L          L'_n += sum_{m>=n} C(m,m-n) <L_m, d^(m-n)>                              (:1383-1412)
//   L2P          a += -sum_n n <L_n, d^(n-1)>                                            (:1550-1578)
//
// How the constants get into the instruction stream.  Every numeric coefficient (factorials, binomials,
// trinomials, double factorials and their products) lives in a `constexpr` TABLE OBJECT whose constructor the
// compiler front end must evaluate (a constexpr variable is constant-initialised by definition): no factorial
// loop, no double-precision arithmetic and no int->double conversion can survive into device code.  The device
// loop nests walk the tables with compile-time indices (full unrolling), so every entry becomes an FFMA
// immediate.  (Round 1 called the constexpr helper FUNCTIONS from the unrolled nests; nvcc does not fold those
// past its unroll budget and left DFMA/DMUL/I2F.F64 loops in the order-4/5 kernels: profiles/r02_sass_fp64.txt.)
//
// Evaluation forms chosen here (not the reference's): the gradient polynomials are evaluated in Horner form in
// dx^2, dy^2; contractions consume source tensors that were multiplied by their trinomial weights once
// ("weighted tuples"), so the inner sums are pure FMA chains.
// The file is host+device so that tests can run the same templates on the CPU against oracle/.
#pragma once

#ifndef __CUDACC__
#define NBCO_HD inline
#define NBCO_UNROLL
#else
#define NBCO_HD __host__ __device__ __forceinline__
#define NBCO_UNROLL _Pragma("unroll")
#endif

namespace nbco { namespace ops {

// ---- index arithmetic (integer constexpr: folded after unrolling) ----
constexpr int sym_elems(int n) { return (n + 1) * (n + 2) / 2; }
constexpr int sym_off(int p) { return p * (p + 1) * (p + 2) / 6; }
constexpr int trl_off(int p) { return p * p; }
constexpr int sym_idx(int x, int z, int n) { return (n * (n + 1) - (n - z) * (n - z + 1)) / 2 + n - x; }
constexpr int trl_idx(int x, int z, int n) { return (z + 1) * n - x; }

// ---- compile-time scalars: ONLY called from the constexpr table constructors below ----
namespace ct {
constexpr double fact(int n) { double r = 1; for (int i = 2; i <= n; ++i) r *= i; return r; }
constexpr double dfact(int n) { double r = 1; for (int i = n; i > 1; i -= 2) r *= i; return r; } // n!!, (-1)!! = 1
constexpr double binom(int n, int k) { return (k < 0 || k > n) ? 0.0 : fact(n) / (fact(k) * fact(n - k)); }
constexpr double trinom(int n, int kx, int kz) { return fact(n) / (fact(kx) * fact(n - kx - kz) * fact(kz)); }
constexpr double pow2(int k) { double r = 1; for (int i = 0; i < k; ++i) r *= 2; return r; }
constexpr double c2(int n, int k) { return fact(n) / (pow2(k) * fact(k) * fact(n - 2 * k)); } // n! / (2^k k! (n-2k)!)
} // namespace ct

// trinomial weights of one order: w[sym_idx(kx,kz,n)] = n! / (kx! ky! kz!)
template <int n>
struct TrinomTab
{
	float w[sym_elems(n)] = {};
	constexpr TrinomTab()
	{
		for (int kz = 0; kz <= n; ++kz)
			for (int kx = 0; kx <= n - kz; ++kx) w[sym_idx(kx, kz, n)] = (float)ct::trinom(n, kx, kz);
	}
};

// (-1)^q / q!
template <int N>
struct SignedInvFactTab
{
	float c[N + 1] = {};
	constexpr SignedInvFactTab() { for (int q = 0; q <= N; ++q) c[q] = (float)(((q & 1) ? -1.0 : 1.0) / ct::fact(q)); }
};

// ---- tensor powers ----
// tuple of tensor powers d^(x) d^(y) d^(z), orders 0..N, symmetric layout; built order by order from the
// previous one (one multiply per entry): entry (x,y,z) of order q is dx * (x-1,y,z), else dy * (0,y-1,z),
// else dz * (0,0,z-1)
template <int N>
NBCO_HD void tensor_pow_tuple(float *pwt /* sym_off(N + 1) */, float dx, float dy, float dz)
{
	pwt[0] = 1.f;
	NBCO_UNROLL
	for (int q = 1; q <= N; ++q)
		NBCO_UNROLL
		for (int z = 0; z <= q; ++z)
			NBCO_UNROLL
			for (int x = q - z; x >= 0; --x)
			{
				const int y = q - x - z;
				float v;
				if (x > 0) v = dx * pwt[sym_off(q - 1) + sym_idx(x - 1, z, q - 1)];
				else if (y > 0) v = dy * pwt[sym_off(q - 1) + sym_idx(0, z, q - 1)];
				else v = dz * pwt[sym_off(q - 1) + sym_idx(0, z - 1, q - 1)];
				pwt[sym_off(q) + sym_idx(x, z, q)] = v;
			}
}

// multiply every order q = 0..N of a symmetric tuple by its trinomial weights (in place)
template <int N, int q = 0>
NBCO_HD void weight_tuple(float *T)
{
	if constexpr (q >= 2) // orders 0 and 1 have unit weights
	{
		constexpr TrinomTab<q> tab{};
		NBCO_UNROLL
		for (int i = 0; i < sym_elems(q); ++i) T[sym_off(q) + i] *= tab.w[i];
	}
	if constexpr (q + 1 <= N) weight_tuple<N, q + 1>(T);
}

// ---- P2M: orders 2..P-1 of a leaf multipole about its centre (dipole == 0, monopole = count) ----
template <int P>
NBCO_HD void p2m_acc(float *M /* sym_off(P) */, float dx, float dy, float dz)
{
	if constexpr (P >= 3)
	{
		constexpr SignedInvFactTab<P - 1> sf{};
		float pwt[sym_off(P)];
		tensor_pow_tuple<P - 1>(pwt, dx, dy, dz);
		NBCO_UNROLL
		for (int q = 2; q <= P - 1; ++q)
			NBCO_UNROLL
			for (int i = 0; i < sym_elems(q); ++i)
				M[sym_off(q) + i] += sf.c[q] * pwt[sym_off(q) + i];
	}
}

// ---- M2M: shift a child tuple (orders 0..P-1, dipole slot zero) by d = new - old centre, orders 2..P-1 ----
// Source-driven form of (:1111-1146): the source term M_s[a,b,c] (order s) lands on the target entry
// [a+k1, b+k2, c+k3] of order s+m with weight C(a+k1,k1) C(b+k2,k2) C(c+k3,k3) s!/(s+m)! d^k.  Same terms as the
// target-driven sum, but every loop bound is a plain triangle.  The weights are tabulated in loop order.
constexpr int m2m_terms(int P)
{
	int t = 0;
	for (int s = 0; s <= P - 1; ++s)
	{
		if (s == 1) continue;
		for (int m = (s >= 2 ? 0 : 2); m <= P - 1 - s; ++m) t += sym_elems(s) * sym_elems(m);
	}
	return t;
}

template <int P>
struct M2MTab
{
	float w[m2m_terms(P) > 0 ? m2m_terms(P) : 1] = {};
	constexpr M2MTab()
	{
		int t = 0;
		for (int s = 0; s <= P - 1; ++s)
		{
			if (s == 1) continue;
			for (int c = 0; c <= s; ++c)
				for (int a = 0; a <= s - c; ++a)
				{
					const int b = s - a - c;
					for (int m = (s >= 2 ? 0 : 2); m <= P - 1 - s; ++m)
						for (int k3 = 0; k3 <= m; ++k3)
							for (int k1 = 0; k1 <= m - k3; ++k1)
							{
								const int k2 = m - k1 - k3;
								w[t++] = (float)(ct::binom(a + k1, k1) * ct::binom(b + k2, k2) * ct::binom(c + k3, k3) * ct::fact(s) / ct::fact(s + m));
							}
				}
		}
	}
};

template <int P>
NBCO_HD void m2m_acc(float *Mout, const float *Min, float dx, float dy, float dz)
{
	if constexpr (P >= 3)
	{
		constexpr M2MTab<P> tab{};
		float pwt[sym_off(P)];
		tensor_pow_tuple<P - 1>(pwt, dx, dy, dz);
		int t = 0; // walks the table in the constructor's loop order: a compile-time constant after unrolling
		NBCO_UNROLL
		for (int s = 0; s <= P - 1; ++s)
		{
			if (s == 1) continue; // dipole of a centre-of-charge expansion is zero
			NBCO_UNROLL
			for (int c = 0; c <= s; ++c)
				NBCO_UNROLL
				for (int a = 0; a <= s - c; ++a)
				{
					const float src = Min[sym_off(s) + sym_idx(a, c, s)];
					NBCO_UNROLL
					for (int m = (s >= 2 ? 0 : 2); m <= P - 1 - s; ++m) // targets of order n = s + m in 2..P-1
						NBCO_UNROLL
						for (int k3 = 0; k3 <= m; ++k3)
							NBCO_UNROLL
							for (int k1 = 0; k1 <= m - k3; ++k1)
							{
								const int n = s + m;
								Mout[sym_off(n) + sym_idx(a + k1, c + k3, n)] += (tab.w[t] * pwt[sym_off(m) + sym_idx(k1, k3, m)]) * src;
								++t;
							}
				}
		}
	}
}

// ---- gradient of 1/r, order m, scaled by r^(m+1): full symmetric layout ----
// Entry (x,y,z), z in {0,1}, is the bivariate polynomial
//     (-1)^m dz^z dx^(x&1) dy^(y&1) sum_{k1 <= X} u^(X-k1) sum_{k2 <= Y} W(k1,k2) v^(Y-k2),   u = dx^2, v = dy^2, X = x/2, Y = y/2
//     W(k1,k2) = (-1)^(k1+k2) (2(m-k1-k2)-1)!! c2(x,k1) c2(y,k2)
// evaluated as Horner chains in v inside a Horner chain in u; the z >= 2 entries follow from the trace relation.
// d = unit-ish direction (d / r with r softened), so the result is dimensionless.
constexpr int grad_terms(int m)
{
	int t = 0;
	for (int z = 0; z <= (m < 1 ? m : 1); ++z)
		for (int x = m - z; x >= 0; --x) t += (x / 2 + 1) * ((m - x - z) / 2 + 1);
	return t;
}

template <int m>
struct GradTab
{
	float w[grad_terms(m)] = {};
	constexpr GradTab()
	{
		int t = 0;
		const double sgn = (m & 1) ? -1.0 : 1.0;
		for (int z = 0; z <= (m < 1 ? m : 1); ++z)
			for (int x = m - z; x >= 0; --x)
			{
				const int y = m - x - z;
				for (int k1 = 0; k1 <= x / 2; ++k1)
					for (int k2 = 0; k2 <= y / 2; ++k2)
						w[t++] = (float)(sgn * (((k1 + k2) & 1) ? -1.0 : 1.0) * ct::dfact(2 * (m - k1 - k2) - 1) * ct::c2(x, k1) * ct::c2(y, k2));
			}
	}
};

// odd[z][x&1][y&1] = dz^z dx^(x&1) dy^(y&1)
struct DirPows { float u, v, odd[2][2][2]; };
NBCO_HD DirPows make_dir_pows(float dx, float dy, float dz)
{
	DirPows p;
	p.u = dx * dx; p.v = dy * dy;
	p.odd[0][0][0] = 1.f; p.odd[0][0][1] = dy; p.odd[0][1][0] = dx; p.odd[0][1][1] = dx * dy;
	p.odd[1][0][0] = dz; p.odd[1][0][1] = dz * dy; p.odd[1][1][0] = dz * dx; p.odd[1][1][1] = dz * p.odd[0][1][1];
	return p;
}

template <int m>
NBCO_HD void grad_scaled(float *g /* sym_elems(m) */, const DirPows &dp)
{
	constexpr GradTab<m> tab{};
	int t = 0;
	NBCO_UNROLL
	for (int z = 0; z <= (m < 1 ? m : 1); ++z)
		NBCO_UNROLL
		for (int x = m - z; x >= 0; --x)
		{
			const int y = m - x - z;
			float outer = 0.f;
			NBCO_UNROLL
			for (int k1 = 0; k1 <= x / 2; ++k1)
			{
				float inner = tab.w[t]; ++t;
				NBCO_UNROLL
				for (int k2 = 1; k2 <= y / 2; ++k2) { inner = inner * dp.v + tab.w[t]; ++t; }
				outer = (k1 == 0) ? inner : outer * dp.u + inner;
			}
			g[sym_idx(x, z, m)] = outer * dp.odd[z][x & 1][y & 1];
		}
	NBCO_UNROLL
	for (int z = 2; z <= m; ++z)
		NBCO_UNROLL
		for (int x = m - z; x >= 0; --x)
			g[sym_idx(x, z, m)] = -g[sym_idx(x + 2, z - 2, m)] - g[sym_idx(x, z - 2, m)];
}

// C[traceless, order nA-nB] += c * <A (order nA, symmetric layout), BW (order nB, symmetric layout, ALREADY
// multiplied by the trinomial weights of order nB)>
template <int nA, int nB>
NBCO_HD void contract_trl_w(float *C, const float *A, const float *BW, float c)
{
	constexpr int nC = nA - nB;
	NBCO_UNROLL
	for (int z = 0; z <= (nC < 1 ? nC : 1); ++z)
		NBCO_UNROLL
		for (int x = nC - z; x >= 0; --x)
		{
			float t = 0.f;
			NBCO_UNROLL
			for (int kz = 0; kz <= nB; ++kz)
				NBCO_UNROLL
				for (int kx = 0; kx <= nB - kz; ++kx)
					t += A[sym_idx(x + kx, z + kz, nA)] * BW[sym_idx(kx, kz, nB)];
			C[trl_idx(x, z, nC)] += c * t;
		}
}

template <int N>
struct InvFactTab
{
	float c[N + 1] = {};
	constexpr InvFactTab() { for (int q = 0; q <= N; ++q) c[q] = (float)(1.0 / ct::fact(q)); }
};

template <int P, int m>
struct M2LStep
{
	// order-m gradient against the multipole orders k = m - n (n = 1..m, k != 1)
	template <int n>
	static NBCO_HD void inner(float *L, const float *MW, const float *g, float cm)
	{
		constexpr int k = m - n;
		constexpr InvFactTab<P> inf{};
		if constexpr (k != 1 && k <= P - 1)
			contract_trl_w<m, k>(L + trl_off(n), g, MW + sym_off(k), cm * inf.c[n]);
		if constexpr (n + 1 <= m)
			inner<n + 1>(L, MW, g, cm);
	}
	static NBCO_HD void run(float *L, const float *MW, const DirPows &dp, float rinv, float rinv_m /* rinv^m */)
	{
		float g[sym_elems(m)];
		grad_scaled<m>(g, dp);
		const float cm = rinv_m * rinv; // 1 / r^(m+1)
		inner<1>(L, MW, g, cm);
		if constexpr (m + 1 <= P)
			M2LStep<P, m + 1>::run(L, MW, dp, rinv, cm);
	}
};

// ---- M2L: L (traceless tuple orders 0..P, slot 0 untouched) += translation of M (symmetric tuple
// orders 0..P-1) over d = c_target - c_source; ux,uy,uz = d / r, rinv = 1 / r, r = sqrt(|d|^2 + eps2).
// MW = M multiplied by the trinomial weights (m2l_weight); the weighted tuple serves both directions of a pair
// only when the SOURCE is the same, so the kernels weight each source tuple once after loading it. ----
template <int P>
NBCO_HD void m2l_weight(float *M /* sym_off(P), in place */) { weight_tuple<P - 1>(M); }

template <int P>
NBCO_HD void m2l_acc_w(float *L, const float *MW, float ux, float uy, float uz, float rinv)
{
	const DirPows dp = make_dir_pows(ux, uy, uz);
	M2LStep<P, 1>::run(L, MW, dp, rinv, rinv);
}

template <int P>
NBCO_HD void m2l_acc(float *L, const float *M, float ux, float uy, float uz, float rinv)
{
	float MW[sym_off(P)];
	NBCO_UNROLL
	for (int i = 0; i < sym_off(P); ++i) MW[i] = M[i];
	m2l_weight<P>(MW);
	m2l_acc_w<P>(L, MW, ux, uy, uz, rinv);
}

// expand the traceless orders 1..P of a local tuple into symmetric layout (S has sym_off(P+1) floats)
template <int P>
NBCO_HD void local_expand(float *S, const float *Ltrl)
{
	NBCO_UNROLL
	for (int q = 1; q <= P; ++q)
	{
		NBCO_UNROLL
		for (int j = 0; j < 2 * q + 1; ++j)
			S[sym_off(q) + j] = Ltrl[trl_off(q) + j];
		NBCO_UNROLL
		for (int z = 2; z <= q; ++z)
			NBCO_UNROLL
			for (int x = q - z; x >= 0; --x)
				S[sym_off(q) + sym_idx(x, z, q)] = -S[sym_off(q) + sym_idx(x + 2, z - 2, q)] - S[sym_off(q) + sym_idx(x, z - 2, q)];
	}
}

template <int N>
struct BinomTab
{
	float c[N + 1][N + 1] = {};
	constexpr BinomTab()
	{
		for (int n = 0; n <= N; ++n)
			for (int k = 0; k <= N; ++k) c[n][k] = (float)ct::binom(n, k);
	}
};

template <int P, int n, int m>
struct L2LStep
{
	static NBCO_HD void run(float *Lout, const float *S, const float *pww /* weighted tensor powers */)
	{
		constexpr BinomTab<P> bt{};
		contract_trl_w<m, m - n>(Lout + trl_off(n), S + sym_off(m), pww + sym_off(m - n), bt.c[m][m - n]);
		if constexpr (m + 1 <= P)
			L2LStep<P, n, m + 1>::run(Lout, S, pww);
		else if constexpr (n + 1 <= P)
			L2LStep<P, n + 1, n + 1>::run(Lout, S, pww);
	}
};

// tensor powers of orders 0..N times their trinomial weights: the terms of the multinomial expansion of (d)^(q)
template <int N>
NBCO_HD void weighted_pow_tuple(float *pww /* sym_off(N + 1) */, float dx, float dy, float dz)
{
	tensor_pow_tuple<N>(pww, dx, dy, dz);
	weight_tuple<N>(pww);
}

// ---- L2L: child tuple (traceless) += shift of the parent tuple given in symmetric layout S ----
template <int P>
NBCO_HD void l2l_acc(float *Lchild, const float *S, float dx, float dy, float dz)
{
	float pww[sym_off(P)]; // orders 0..P-1
	weighted_pow_tuple<P - 1>(pww, dx, dy, dz);
	L2LStep<P, 1, 1>::run(Lchild, S, pww);
}

template <int P, int n>
struct L2PStep
{
	static NBCO_HD void run(float *f, const float *S, const float *pww)
	{
		contract_trl_w<n, n - 1>(f, S + sym_off(n), pww + sym_off(n - 1), (float)n);
		if constexpr (n + 1 <= P)
			L2PStep<P, n + 1>::run(f, S, pww);
	}
};

// ---- L2P: field at offset d from the leaf centre; f[3] receives -sum_n n <L_n, d^(n-1)> ----
template <int P>
NBCO_HD void l2p_field(float *f, const float *S, float dx, float dy, float dz)
{
	float pww[sym_off(P)];
	weighted_pow_tuple<P - 1>(pww, dx, dy, dz);
	float t[3] = {0.f, 0.f, 0.f};
	L2PStep<P, 1>::run(t, S, pww);
	f[0] = -t[0]; f[1] = -t[1]; f[2] = -t[2];
}

}} // namespace nbco::ops
