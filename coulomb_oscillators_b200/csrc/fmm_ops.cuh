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
// Everything is a constexpr-indexed loop nest: with P a template parameter nvcc unrolls the nests
// completely, folds every coefficient into an immediate and keeps the tensors in registers.
// The file is host+device so that tests can run the same templates on the CPU against oracle/.
#pragma once

#ifndef __CUDACC__
#define NBCO_HD
#define NBCO_UNROLL
#else
#define NBCO_HD __host__ __device__ __forceinline__
#define NBCO_UNROLL _Pragma("unroll")
#endif

namespace nbco { namespace ops {

// ---- integer helpers (all constexpr: folded after unrolling) ----
constexpr int sym_elems(int n) { return (n + 1) * (n + 2) / 2; }
constexpr int sym_off(int p) { return p * (p + 1) * (p + 2) / 6; }
constexpr int trl_off(int p) { return p * p; }
constexpr int sym_idx(int x, int z, int n) { return (n * (n + 1) - (n - z) * (n - z + 1)) / 2 + n - x; }
constexpr int trl_idx(int x, int z, int n) { return (z + 1) * n - x; }

constexpr double cfact(int n) { double r = 1; for (int i = 2; i <= n; ++i) r *= i; return r; }
constexpr double codfact(int n) { double r = 1; for (int i = n; i > 1; i -= 2) r *= i; return r; }
constexpr double cbinom(int n, int k) { return (k < 0 || k > n) ? 0.0 : cfact(n) / (cfact(k) * cfact(n - k)); }
constexpr double ctrinom(int n, int kx, int kz) { return cfact(n) / (cfact(kx) * cfact(n - kx - kz) * cfact(kz)); }
constexpr double cpow2(int k) { double r = 1; for (int i = 0; i < k; ++i) r *= 2; return r; }
constexpr double ccoeff13(int n, int m) { return ((m & 1) ? -1.0 : 1.0) * codfact(2 * (n - m) - 1); }
constexpr double ccoeff2(int n, int k) { return cfact(n) / (cpow2(k) * cfact(k) * cfact(n - 2 * k)); }

// powers d^0..d^N of the three components
template <int N>
struct Pow3
{
	float x[N + 1], y[N + 1], z[N + 1];
	NBCO_HD Pow3(float dx, float dy, float dz)
	{
		x[0] = y[0] = z[0] = 1.f;
		NBCO_UNROLL
		for (int i = 1; i <= N; ++i) { x[i] = x[i-1] * dx; y[i] = y[i-1] * dy; z[i] = z[i-1] * dz; }
	}
};

// ---- P2M: orders 2..P-1 of a leaf multipole about its centre (dipole == 0, monopole = count) ----
template <int P>
NBCO_HD void p2m_acc(float *M /* sym_off(P) */, float dx, float dy, float dz)
{
	if constexpr (P >= 3)
	{
		Pow3<P - 1> pw(dx, dy, dz);
		NBCO_UNROLL
		for (int q = 2; q <= P - 1; ++q)
		{
			const float C = (float)(((q & 1) ? -1.0 : 1.0) / cfact(q));
			NBCO_UNROLL
			for (int z = 0; z <= q; ++z)
				NBCO_UNROLL
				for (int x = q - z; x >= 0; --x)
					M[sym_off(q) + sym_idx(x, z, q)] += C * pw.x[x] * pw.y[q - x - z] * pw.z[z];
		}
	}
}

// ---- M2M: shift a child tuple (orders 0..P-1, dipole slot zero) by d = new - old centre, orders 2..P-1 ----
// Source-driven form of (:1111-1146): the source term M_s[a,b,c] (order s) lands on the target entry
// [a+k1, b+k2, c+k3] of order s+m with weight C(a+k1,k1) C(b+k2,k2) C(c+k3,k3) s!/(s+m)! d^k.  Same terms as the
// target-driven sum, but every loop bound is a plain triangle (the max/min bounds of the target-driven nest made
// nvcc's unroller explode: order 4 did not compile in 15 minutes).
template <int P>
NBCO_HD void m2m_acc(float *Mout, const float *Min, float dx, float dy, float dz)
{
	if constexpr (P >= 3)
	{
		Pow3<P - 1> pw(dx, dy, dz);
		NBCO_UNROLL
		for (int s = 0; s <= P - 1; ++s)
		{
			if (s == 1) continue; // dipole of a centre-of-charge expansion is zero
			NBCO_UNROLL
			for (int c = 0; c <= s; ++c)
				NBCO_UNROLL
				for (int a = 0; a <= s - c; ++a)
				{
					const int b = s - a - c;
					const float src = Min[sym_off(s) + sym_idx(a, c, s)];
					NBCO_UNROLL
					for (int m = (s >= 2 ? 0 : 2); m <= P - 1 - s; ++m) // targets of order n = s + m in 2..P-1
						NBCO_UNROLL
						for (int k3 = 0; k3 <= m; ++k3)
							NBCO_UNROLL
							for (int k1 = 0; k1 <= m - k3; ++k1)
							{
								const int k2 = m - k1 - k3, n = s + m;
								const float w = (float)(cbinom(a + k1, k1) * cbinom(b + k2, k2) * cbinom(c + k3, k3) * cfact(s) / cfact(n));
								Mout[sym_off(n) + sym_idx(a + k1, c + k3, n)] += w * pw.x[k1] * pw.y[k2] * pw.z[k3] * src;
							}
				}
		}
	}
}

// ---- gradient of 1/r, order m, scaled by r^(m+1): full symmetric layout ----
// d = unit-ish direction (d / r with r softened), so the result is dimensionless.
template <int m, int PW>
NBCO_HD void grad_scaled(float *g /* sym_elems(m) */, const Pow3<PW> &pw)
{
	const float sgn = (m & 1) ? -1.f : 1.f;
	NBCO_UNROLL
	for (int z = 0; z <= 1; ++z)
		NBCO_UNROLL
		for (int x = m - z; x >= 0; --x)
		{
			const int y = m - x - z;
			float t1 = 0.f;
			NBCO_UNROLL
			for (int k1 = 0; k1 <= x / 2; ++k1)
			{
				float t2 = 0.f;
				NBCO_UNROLL
				for (int k2 = 0; k2 <= y / 2; ++k2)
					t2 += (float)(ccoeff13(m, k1 + k2) * ccoeff2(y, k2)) * pw.y[y - 2 * k2];
				t1 += t2 * (float)ccoeff2(x, k1) * pw.x[x - 2 * k1];
			}
			g[sym_idx(x, z, m)] = sgn * t1 * pw.z[z];
		}
	NBCO_UNROLL
	for (int z = 2; z <= m; ++z)
		NBCO_UNROLL
		for (int x = m - z; x >= 0; --x)
			g[sym_idx(x, z, m)] = -g[sym_idx(x + 2, z - 2, m)] - g[sym_idx(x, z - 2, m)];
}

// C[traceless, order nA-nB] += c * <A (order nA, symmetric layout), B (order nB, symmetric layout)>
template <int nA, int nB>
NBCO_HD void contract_trl_ma(float *C, const float *A, const float *B, float c)
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
					t += (float)ctrinom(nB, kx, kz) * A[sym_idx(x + kx, z + kz, nA)] * B[sym_idx(kx, kz, nB)];
			C[trl_idx(x, z, nC)] += c * t;
		}
}

template <int P, int m>
struct M2LStep
{
	template <int n>
	static NBCO_HD void inner(float *L, const float *M, const float *g, float cm)
	{
		constexpr int k = m - n;
		if constexpr (k != 1)
			contract_trl_ma<m, k>(L + trl_off(n), g, M + sym_off(k), cm * (float)(1.0 / cfact(n)));
		if constexpr (n + 1 <= m)
			inner<n + 1>(L, M, g, cm);
	}
	static NBCO_HD void run(float *L, const float *M, const Pow3<P> &pw, float rinv, float rinv_m /* rinv^m */)
	{
		float g[sym_elems(m)];
		grad_scaled<m, P>(g, pw);
		const float cm = rinv_m * rinv; // 1 / r^(m+1)
		inner<1>(L, M, g, cm);
		if constexpr (m + 1 <= P)
			M2LStep<P, m + 1>::run(L, M, pw, rinv, cm);
	}
};

// ---- M2L: L (traceless tuple orders 0..P, slot 0 untouched) += translation of M (symmetric tuple
// orders 0..P-1) over d = c_target - c_source; ux,uy,uz = d / r, rinv = 1 / r, r = sqrt(|d|^2 + eps2) ----
template <int P>
NBCO_HD void m2l_acc(float *L, const float *M, float ux, float uy, float uz, float rinv)
{
	Pow3<P> pw(ux, uy, uz);
	M2LStep<P, 1>::run(L, M, pw, rinv, rinv);
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

template <int P, int n, int m>
struct L2LStep
{
	static NBCO_HD void run(float *Lout, const float *S, const float *pwt /* tensor powers tuple */)
	{
		contract_trl_ma<m, m - n>(Lout + trl_off(n), S + sym_off(m), pwt + sym_off(m - n), (float)cbinom(m, m - n));
		if constexpr (m + 1 <= P)
			L2LStep<P, n, m + 1>::run(Lout, S, pwt);
		else if constexpr (n + 1 <= P)
			L2LStep<P, n + 1, n + 1>::run(Lout, S, pwt);
	}
};

// tuple of tensor powers d^(x) d^(y) d^(z), orders 0..N, symmetric layout
template <int N>
NBCO_HD void tensor_pow_tuple(float *pwt, float dx, float dy, float dz)
{
	Pow3<N> pw(dx, dy, dz);
	NBCO_UNROLL
	for (int q = 0; q <= N; ++q)
		NBCO_UNROLL
		for (int z = 0; z <= q; ++z)
			NBCO_UNROLL
			for (int x = q - z; x >= 0; --x)
				pwt[sym_off(q) + sym_idx(x, z, q)] = pw.x[x] * pw.y[q - x - z] * pw.z[z];
}

// ---- L2L: child tuple (traceless) += shift of the parent tuple given in symmetric layout S ----
template <int P>
NBCO_HD void l2l_acc(float *Lchild, const float *S, float dx, float dy, float dz)
{
	float pwt[sym_off(P)]; // orders 0..P-1
	tensor_pow_tuple<P - 1>(pwt, dx, dy, dz);
	L2LStep<P, 1, 1>::run(Lchild, S, pwt);
}

template <int P, int n>
struct L2PStep
{
	static NBCO_HD void run(float *f, const float *S, const float *pwt)
	{
		contract_trl_ma<n, n - 1>(f, S + sym_off(n), pwt + sym_off(n - 1), (float)n);
		if constexpr (n + 1 <= P)
			L2PStep<P, n + 1>::run(f, S, pwt);
	}
};

// ---- L2P: field at offset d from the leaf centre; f[3] receives -sum_n n <L_n, d^(n-1)> ----
template <int P>
NBCO_HD void l2p_field(float *f, const float *S, float dx, float dy, float dz)
{
	float pwt[sym_off(P)];
	tensor_pow_tuple<P - 1>(pwt, dx, dy, dz);
	float t[3] = {0.f, 0.f, 0.f};
	L2PStep<P, 1>::run(t, S, pwt);
	f[0] = -t[0]; f[1] = -t[1]; f[2] = -t[2];
}

}} // namespace nbco::ops
