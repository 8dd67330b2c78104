/* nbco_oracle2d.c -- CPU restatement of the reference's 2D fp64 path (uniform quadtree FMM, direct
 * sum, integrators).  TEST INFRASTRUCTURE ONLY (see nbco_oracle.c for the rules).
 *
 * Follows fmm_cart_cpu (reference Simulation/fmm_cart.cuh:546-680) and the 2D operator algebra of
 * fmm_cart_base.cuh; SCAL = double, VEC = double2.  Pinned against the unmodified reference compiled with
 * -DSCAL=double -DDIM=2 behind oracle/ref_harness2d.cu (tests/test_oracle2d.py) and the fixtures generated
 * from it (tests/golden/fmm2d_*.npz).  The only freedom is the order of particles inside one grid cell
 * (the reference's CPU sort is unstable, its GPU sort stable): a stable sort by cell key is used here. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define O2_MAX_ORDER 12

static double fact2(int n) { double r = 1; for (int i = 2; i <= n; ++i) r *= i; return r; }
static double edfact(int n) { double r = 1; for (int i = n; i > 1; i -= 2) r *= i; return r; } /* n!!, n even, 0!! = 1 */
static double binom2(int n, int k) { return (k < 0 || k > n) ? 0 : fact2(n) / (fact2(k) * fact2(n - k)); }
static double ipowd(double b, int e) { double r = 1; if (e < 0) { b = 1 / b; e = -e; } for (int i = 0; i < e; ++i) r *= b; return r; }
static double coeff1(int n, int m) { return ((m & 1) ? -1.0 : 1.0) * edfact(2 * (n - m) - 2); }            /* fmm_cart_base.cuh:22-31 */
static double coeff2d(int n, int m) { return fact2(n) / (ldexp(1.0, m) * fact2(m) * fact2(n - 2 * m)); }     /* :33-40 */
static int sym_off2(int p) { return p * (p + 1) / 2; }                                                        /* :111 */
static int trl_off2(int p) { return p == 0 ? 0 : 2 * p - 1; }                                                 /* :116 */

int orc2_sym_off(int p) { return sym_off2(p); }
int orc2_trl_off(int p) { return trl_off2(p); }

/* -nabla^n log r, traceless form: entries 0,1 (:373-395); d = unit vector, r = distance */
static void gradient2(double *g, int n, const double d[2], double r)
{
	if (n == 0) { g[0] = -log(r); return; }
	double C = ((n & 1) ? -1.0 : 1.0) * ipowd(r, -n);
	for (int i = 0; i <= 1; ++i)
	{
		int j = n - i;
		double t = 0;
		for (int m = 0; m <= j / 2; ++m) t += coeff1(n, m) * coeff2d(j, m) * ipowd(d[0], j - 2 * m);
		g[i] = C * t * ipowd(d[1], i);
	}
}

/* traceless power of a unit vector scaled by r^n (:434-451) */
static void tracelesspow2(double *pw, int n, const double d[2], double r)
{
	if (n == 0) { pw[0] = 1; pw[1] = 0; return; }
	double C = ipowd(r, n) / edfact(2 * n - 2);
	for (int i = 0; i <= 1; ++i)
	{
		int j = n - i;
		double t = 0;
		for (int m = 0; m <= j / 2; ++m) t += coeff1(n, m) * coeff2d(j, m) * ipowd(d[0], j - 2 * m);
		pw[i] = C * t * ipowd(d[1], i);
	}
}

/* C (order nA-nB, 2 entries) += c <A (traceless, order nA), B (traceless, order nB)> (:234-261) */
static void contract_trl2(double *C, const double *A, const double *B, double c, int nA, int nB)
{
	if (nA < nB) { const double *t = A; A = B; B = t; int k = nA; nA = nB; nB = k; }
	int nC = nA - nB;
	if (nB >= 1)
	{
		double t = c * ldexp(1.0, nB - 1);
		C[0] += t * (A[0] * B[0] + A[1] * B[1]);
		if (nC >= 1) C[1] += t * (A[1] * B[0] - A[0] * B[1]);
	}
	else
	{
		C[0] += c * A[0] * B[0];
		if (nC >= 1) C[1] += c * A[1] * B[0];
	}
}

static void p2m2(double *M, int q, const double d[2])
{
	double C = ((q & 1) ? -1.0 : 1.0) / fact2(q);
	for (int i = 0; i <= q; ++i) M[i] += C * ipowd(d[1], i) * ipowd(d[0], q - i);
}

static void m2m2(double *Mout, const double *Mtuple, int n, const double d[2])
/* :527-551 */
{
	for (int i = 0; i <= n; ++i)
	{
		int j = n - i;
		double t = 0;
		for (int m = 0; m <= n; ++m)
		{
			const double *Mo = Mtuple + sym_off2(n - m);
			double c = 0;
			int lo = m - j > 0 ? m - j : 0, hi = i < m ? i : m;
			for (int k = lo; k <= hi; ++k)
			{
				int l = m - k;
				c += binom2(i, k) * binom2(j, l) * ipowd(d[1], k) * ipowd(d[0], l) * Mo[i - k];
			}
			t += c * (fact2(n - m) / fact2(n));
		}
		Mout[i] += t;
	}
}

static void m2l2(double *Ltuple, const double *Mtuple, int p, double dx, double dy, double r2)
/* static_m2l_acc<1> / m2l_acc with minm = 1 (:691-713, 757-798) */
{
	double g[O2_MAX_ORDER + 3];
	double r = sqrt(r2), d[2] = {dx / r, dy / r};
	for (int m = 1; m <= 2 * p; ++m)
	{
		gradient2(g, m, d, r);
		int top = m < p + 1 ? m : p + 1;
		for (int i = 2; i <= top; ++i) g[i] = -g[i - 2];
		for (int n = (m - p > 0 ? m - p : 0); n <= (p < m ? p : m); ++n)
		{
			int mn = m - n;
			double C = 1.0 / fact2(n);
			const double *M = Mtuple + sym_off2(mn);
			double *Ln = Ltuple + trl_off2(n);
			for (int i = 0; i <= (n < 1 ? n : 1); ++i)
			{
				double t = 0;
				for (int k = 0; k <= mn; ++k) t += binom2(mn, k) * g[i + k] * M[k];
				Ln[i] += C * t;
			}
		}
	}
}

static void l2l2(double *Lout /* order q */, const double *Ltuple, int q, int p, const double d[2], double r)
/* l2l_traceless_acc (:912-927) */
{
	double pw[2];
	for (int m = q; m <= p; ++m)
	{
		tracelesspow2(pw, m - q, d, r);
		contract_trl2(Lout, Ltuple + trl_off2(m), pw, binom2(m, m - q), m, m - q);
	}
}

static void l2p2(double f[2], const double *Ltuple, int p, const double d[2], double r)
/* l2p_traceless_field (:1002-1018) */
{
	double pw[2], t[2] = {0, 0};
	for (int n = 1; n <= p; ++n)
	{
		tracelesspow2(pw, n - 1, d, r);
		contract_trl2(t, Ltuple + trl_off2(n), pw, (double)n, n, n - 1);
	}
	f[0] = -t[0]; f[1] = -t[1];
}

static int tbeg(int l) { return ((1 << (2 * l)) - 1) / 3; }

int orc2_levels(int n, int order, double dens)
/* fmm_cart.cuh:562-564 */
{
	double s = order * sqrt((double)order);
	int L = (int)round(log2(dens * (double)n / s) / 2);
	return L < 2 ? 2 : L;
}

typedef struct { int key, idx; } kv2_t;
static void stable_sort_kv2(kv2_t *a, kv2_t *tmp, int n)
{
	for (int w = 1; w < n; w *= 2)
	{
		for (int lo = 0; lo < n; lo += 2 * w)
		{
			int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
			int i = lo, j = mid, k = lo;
			while (i < mid && j < hi) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
			while (i < mid) tmp[k++] = a[i++];
			while (j < hi) tmp[k++] = a[j++];
		}
		memcpy(a, tmp, sizeof(kv2_t) * (size_t)n);
	}
}

/* fmm_cart_cpu.  pos, vel (may be NULL), acc: n double2; pos and vel are permuted into cell order.
 * Outputs (may be NULL): perm[n]; center (2/node), mpole (offM/node), local (offL/node), mult, index per node
 * (ntot = (4^(L+1)-1)/3 nodes, levels 0..L, only 2..L are used like in the reference). */
int orc2_fmm(double *pos, double *vel, double *acc, int n, const double *param, int order, int radius, double eps2,
             double dens, int coll, int *perm_out, double *center_out, double *mpole_out, double *local_out,
             int *mult_out, int *index_out)
{
	if (order < 1 || order > O2_MAX_ORDER) return -1;
	const int p = order, L = orc2_levels(n, order, dens), side = 1 << L;
	const int ntot = ((1 << (2 * (L + 1))) - 1) / 3, offM = sym_off2(p + 1), offL = trl_off2(p + 1);
	double *center = calloc((size_t)ntot, 16), *mpole = calloc((size_t)ntot * offM, 8), *local = calloc((size_t)ntot * offL, 8);
	int *mult = calloc((size_t)ntot, 4), *index = calloc((size_t)ntot + 1, 4);
	kv2_t *kv = malloc(sizeof(kv2_t) * (size_t)n), *tmp = malloc(sizeof(kv2_t) * (size_t)n);
	double *pt = malloc(16 * (size_t)n);

	double mn[2] = {pos[0], pos[1]}, mx[2] = {pos[0], pos[1]};
	for (int i = 1; i < n; ++i)
		for (int k = 0; k < 2; ++k) { mn[k] = fmin(mn[k], pos[2 * i + k]); mx[k] = fmax(mx[k], pos[2 * i + k]); }
	double delta = fmax(mx[0] - mn[0], mx[1] - mn[1]) / (double)side, eps = sqrt(eps2);
	if (delta < eps) delta = eps;
	const double rdelta = 1.0 / delta;
	for (int i = 0; i < n; ++i)
	{
		/* evalKeys (appel.cuh:44-55): truncate, clip, row-major with x major */
		int ix = (int)((pos[2 * i] - mn[0]) * rdelta), iy = (int)((pos[2 * i + 1] - mn[1]) * rdelta);
		ix = ix < 0 ? 0 : (ix > side - 1 ? side - 1 : ix);
		iy = iy < 0 ? 0 : (iy > side - 1 ? side - 1 : iy);
		kv[i].key = ix * side + iy; kv[i].idx = i;
	}
	stable_sort_kv2(kv, tmp, n);
	for (int i = 0; i < n; ++i) memcpy(pt + 2 * (size_t)i, pos + 2 * (size_t)kv[i].idx, 16);
	memcpy(pos, pt, 16 * (size_t)n);
	if (vel)
	{
		for (int i = 0; i < n; ++i) memcpy(pt + 2 * (size_t)i, vel + 2 * (size_t)kv[i].idx, 16);
		memcpy(vel, pt, 16 * (size_t)n);
	}
	if (perm_out) for (int i = 0; i < n; ++i) perm_out[i] = kv[i].idx;

	const int begL = tbeg(L), m = side * side;
	/* indexLeaves / multLeaves (appel.cuh:141-212): first particle of every cell */
	{
		int c = 0;
		for (int i = 0; i < n; ++i) { while (c <= kv[i].key) index[begL + c++] = i; }
		while (c < m) index[begL + c++] = n;
		for (int i = 0; i < m; ++i) mult[begL + i] = (i == m - 1 ? n : index[begL + i + 1]) - index[begL + i];
	}
	for (int i = 0; i < m; ++i)
	{
		/* centerLeaves (appel.cuh:226-243) + P2M (fmm_cart.cuh:68-98) */
		int c = begL + i, ml = mult[c];
		const double *pi = pos + 2 * (size_t)index[c];
		double t[2] = {0, 0};
		if (ml > 0) { for (int j = 0; j < ml; ++j) { t[0] += pi[2 * j]; t[1] += pi[2 * j + 1]; } t[0] /= (double)ml; t[1] /= (double)ml; }
		center[2 * c] = t[0]; center[2 * c + 1] = t[1];
		double *M = mpole + (size_t)c * offM;
		M[0] = (double)ml;
		if (p >= 2)
			for (int j = 0; j < ml; ++j)
			{
				double d[2] = {pi[2 * j] - t[0], pi[2 * j + 1] - t[1]};
				for (int q = 2; q <= p; ++q) p2m2(M + sym_off2(q), q, d);
			}
	}
	/* M2M (fmm_cart.cuh:115-187), levels L-1 .. 2 */
	for (int l = L - 1; l >= 2; --l)
	{
		int sl = 1 << l, slp = 1 << (l + 1), beg = tbeg(l), begp = tbeg(l + 1);
		for (int ij0 = 0; ij0 < sl * sl; ++ij0)
		{
			int i = ij0 / sl, j = ij0 - i * sl, ij = beg + ij0, ijp = begp + 2 * (i * slp + j);
			int ch[4] = {ijp, ijp + 1, ijp + slp, ijp + slp + 1};
			int ml = 0;
			for (int k = 0; k < 4; ++k) ml += mult[ch[k]];
			double co[2] = {0, 0};
			if (ml > 0)
			{
				for (int k = 0; k < 4; ++k) { co[0] += (double)mult[ch[k]] * center[2 * ch[k]]; co[1] += (double)mult[ch[k]] * center[2 * ch[k] + 1]; }
				co[0] /= (double)ml; co[1] /= (double)ml;
				double *M = mpole + (size_t)ij * offM;
				if (p >= 2)
					for (int k = 0; k < 4; ++k)
					{
						double d[2] = {co[0] - center[2 * ch[k]], co[1] - center[2 * ch[k] + 1]};
						for (int q = 2; q <= p; ++q) m2m2(M + sym_off2(q), mpole + (size_t)ch[k] * offM, q, d);
					}
				M[0] = (double)ml;
			}
			center[2 * ij] = co[0]; center[2 * ij + 1] = co[1];
			mult[ij] = ml;
		}
	}
	/* near field p2p2 (appel.cuh:260-303): writes acc */
	if (coll)
	{
		for (int ij = 0; ij < m; ++ij)
		{
			int i = ij / side, j = ij - i * side;
			int kmin = i - radius > 0 ? i - radius : 0, kmax = i + radius < side - 1 ? i + radius : side - 1;
			int lmin = j - radius > 0 ? j - radius : 0, lmax = j + radius < side - 1 ? j + radius : side - 1;
			int i1 = index[begL + ij];
			for (int h = 0; h < mult[begL + ij]; ++h)
			{
				double ax = 0, ay = 0, px = pos[2 * (size_t)(i1 + h)], py = pos[2 * (size_t)(i1 + h) + 1];
				for (int k = kmin; k <= kmax; ++k)
				{
					int kl = k * side + lmin, iT = index[begL + kl], mT = 0;
					for (int l = 0; l <= lmax - lmin; ++l) mT += mult[begL + kl + l];
					for (int g = 0; g < mT; ++g)
					{
						double dx = px - pos[2 * (size_t)(iT + g)], dy = py - pos[2 * (size_t)(iT + g) + 1];
						double inv = 1.0 / (dx * dx + dy * dy + eps2);
						ax = fma(inv, dx, ax); ay = fma(inv, dy, ay);
					}
				}
				acc[2 * (size_t)(i1 + h)] = ax; acc[2 * (size_t)(i1 + h) + 1] = ay;
			}
		}
	}
	else memset(acc, 0, 16 * (size_t)n);
	/* M2L (fmm_cart.cuh:214-262), levels L .. 2 */
	for (int l = L; l >= 2; --l)
	{
		int sl = 1 << l, beg = tbeg(l);
		for (int ij = 0; ij < sl * sl; ++ij)
		{
			int i = ij / sl, j = ij - i * sl, ij1 = beg + ij;
			if (mult[ij1] <= 0) continue;
			int im = (i / 2) * 2, jm = (j / 2) * 2;
			int kmin = im - 2 * radius > 0 ? im - 2 * radius : 0, kmax = im + 2 * radius + 1 < sl - 1 ? im + 2 * radius + 1 : sl - 1;
			int gmin = jm - 2 * radius > 0 ? jm - 2 * radius : 0, gmax = jm + 2 * radius + 1 < sl - 1 ? jm + 2 * radius + 1 : sl - 1;
			for (int k = kmin; k <= kmax; ++k)
				for (int g = gmin; g <= gmax; ++g)
				{
					if (!(k > i + radius || k < i - radius || g > j + radius || g < j - radius)) continue;
					int ij2 = beg + k * sl + g;
					double dx = center[2 * ij1] - center[2 * ij2], dy = center[2 * ij1 + 1] - center[2 * ij2 + 1];
					m2l2(local + (size_t)ij1 * offL, mpole + (size_t)ij2 * offM, p, dx, dy, dx * dx + dy * dy + eps2);
				}
		}
	}
	/* L2L (fmm_cart.cuh:288-334), levels 2 .. L-1 */
	for (int l = 2; l <= L - 1; ++l)
	{
		int sl = 1 << l, slp = 1 << (l + 1), beg = tbeg(l), begp = tbeg(l + 1);
		for (int ij0 = 0; ij0 < sl * sl; ++ij0)
		{
			int i = ij0 / sl, j = ij0 - i * sl, ij = beg + ij0, ijp = begp + 2 * (i * slp + j);
			int ch[4] = {ijp, ijp + 1, ijp + slp, ijp + slp + 1};
			for (int k = 0; k < 4; ++k)
			{
				double d[2] = {center[2 * ch[k]] - center[2 * ij], center[2 * ch[k] + 1] - center[2 * ij + 1]};
				double r = sqrt(d[0] * d[0] + d[1] * d[1]);
				if (r != 0) { d[0] /= r; d[1] /= r; }
				for (int q = 0; q <= p; ++q)
					l2l2(local + (size_t)ch[k] * offL + trl_off2(q), local + (size_t)ij * offL, q, p, d, r);
			}
		}
	}
	/* L2P (fmm_cart.cuh:353-378) + rescale */
	for (int i = 0; i < m; ++i)
	{
		int c = begL + i;
		for (int j = 0; j < mult[c]; ++j)
		{
			size_t q = (size_t)(index[c] + j);
			double d[2] = {pos[2 * q] - center[2 * c], pos[2 * q + 1] - center[2 * c + 1]};
			double r = sqrt(d[0] * d[0] + d[1] * d[1]), f[2];
			if (r != 0) { d[0] /= r; d[1] /= r; }
			l2p2(f, local + (size_t)c * offL, p, d, r);
			acc[2 * q] += f[0]; acc[2 * q + 1] += f[1];
		}
	}
	if (param) for (size_t i = 0; i < 2 * (size_t)n; ++i) acc[i] *= param[0];

	if (center_out) memcpy(center_out, center, 16 * (size_t)ntot);
	if (mpole_out) memcpy(mpole_out, mpole, 8 * (size_t)ntot * offM);
	if (local_out) memcpy(local_out, local, 8 * (size_t)ntot * offL);
	if (mult_out) memcpy(mult_out, mult, 4 * (size_t)ntot);
	if (index_out) memcpy(index_out, index, 4 * (size_t)ntot);
	free(center); free(mpole); free(local); free(mult); free(index); free(kv); free(tmp); free(pt);
	return L;
}

void orc2_direct(const double *p, double *a, int n, const double *param, double eps2)
/* direct2_core (direct.cuh:140-164), 2D kernel fma(invDist2, d, a) */
{
	double k = param ? param[0] : 1.0;
	for (int i = 0; i < n; ++i)
	{
		double ax = 0, ay = 0;
		for (int j = 0; j < n; ++j)
		{
			double dx = p[2 * i] - p[2 * j], dy = p[2 * i + 1] - p[2 * j + 1];
			double inv = 1.0 / (dx * dx + dy * dy + eps2);
			ax = fma(inv, dx, ax); ay = fma(inv, dy, ay);
		}
		a[2 * i] = k * ax; a[2 * i + 1] = k * ay;
	}
}

void orc2_add_elastic(const double *p, double *a, int n, const double *k2)
{
	for (int i = 0; i < n; ++i) for (int q = 0; q < 2; ++q) a[2 * i + q] -= p[2 * i + q] * (k2 ? k2[q] : 1.0);
}

void orc2_step(double *b, const double *a, double ds, int n)
{
	for (size_t i = 0; i < 2 * (size_t)n; ++i) b[i] += a[i] * ds;
}

/* evaluator: 0 direct2, 1 fmm_cart, 2 direct2 + elastic, 3 fmm_cart + elastic; buf = [pos|vel|acc] */
int orc2_eval(int evaluator, double *buf, int n, const double *param, int order, int radius, double eps2, double dens, int coll)
{
	double *pos = buf, *vel = buf + 2 * (size_t)n, *acc = buf + 4 * (size_t)n;
	int s = 0;
	if (evaluator == 0 || evaluator == 2) orc2_direct(pos, acc, n, param, eps2);
	else s = orc2_fmm(pos, vel, acc, n, param, order, radius, eps2, dens, coll, NULL, NULL, NULL, NULL, NULL, NULL) < 0 ? -1 : 0;
	if (evaluator >= 2) orc2_add_elastic(pos, acc, n, param ? param + 2 : NULL);
	return s;
}

int orc2_integrate(int scheme, int evaluator, double *buf, int n, const double *param, double dt_, long long nsteps,
                   int order, int radius, double eps2, double dens, int coll)
/* integrator.cuh:32-167 with SCAL = double */
{
	double *pos = buf, *vel = buf + 2 * (size_t)n, *acc = buf + 4 * (size_t)n;
	const long double dt = dt_;
#define K(cf) orc2_step(vel, acc, (double)(cf), n)
#define D(cf) orc2_step(pos, vel, (double)(cf), n)
#define F() do { if (orc2_eval(evaluator, buf, n, param, order, radius, eps2, dens, coll)) return -1; } while (0)
	for (long long s = 0; s < nsteps; ++s)
		switch (scheme)
		{
			case 0: K(dt); D(dt); F(); break;
			case 1: K(dt * 0.5L); D(dt); F(); K(dt * 0.5L); break;
			case 2:
			{
				const long double th = 1.3512071919596576340476878089715L;
				D(dt * th / 2); F(); K(dt * th); D(dt * (1 - th) / 2); F(); K(dt * (1 - 2 * th)); D(dt * (1 - th) / 2); F(); K(dt * th); D(dt * th / 2);
				break;
			}
			case 3:
			{
				const long double xi = +0.1786178958448091E+00L, la = -0.2123418310626054E+00L, ch = -0.6626458266981849E-01L;
				D(dt * xi); F(); K(dt * (1 - 2 * la) / 2); D(dt * ch); F(); K(dt * la); D(dt * (1 - 2 * (ch + xi))); F();
				K(dt * la); D(dt * ch); F(); K(dt * (1 - 2 * la) / 2); D(dt * xi);
				break;
			}
			default: return -1;
		}
#undef K
#undef D
#undef F
	return 0;
}

/* energy of the 2D system: pair term -1/2 log(d^2 + eps2) */
void orc2_energy(const double *buf, int n, const double *param, double eps2, double *out)
{
	const double *x = buf, *v = buf + 2 * (size_t)n;
	double ke = 0, el = 0, pe = 0;
	for (int i = 0; i < n; ++i)
		for (int q = 0; q < 2; ++q)
		{
			ke += 0.5 * v[2 * i + q] * v[2 * i + q];
			el += 0.5 * (param ? param[2 + q] : 1.0) * x[2 * i + q] * x[2 * i + q];
		}
	for (int i = 0; i < n; ++i)
		for (int j = i + 1; j < n; ++j)
		{
			double dx = x[2 * i] - x[2 * j], dy = x[2 * i + 1] - x[2 * j + 1];
			pe += -0.5 * log(dx * dx + dy * dy + eps2);
		}
	out[0] = ke; out[1] = el; out[2] = (param ? param[0] : 1.0) * pe;
}
