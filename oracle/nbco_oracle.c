/* nbco_oracle.c -- CPU restatement of the reference's force-evaluation / time-stepping path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this file; the product (libnbco.so) never does and has no CPU path.
 *
 * Plain C11, single-threaded, written from the algorithm description in SURVEY.md section 2.4/2.5 and
 * the reference sources cited per function (paths relative to /root/reference/Simulation).  It
 * restates the reference's CPU path (fmm_cart3_kdtree_cpu, fmm_cart3_kdtree.cuh:1773-1929) with
 * three declared choices where the reference is undefined or has two behaviours:
 *   (1) ties between equal fp32 keys: every level is a STABLE sort (the reference's level-0 CUB
 *       sort is stable, deeper levels use unstable sorts: :1348-1351,1370-1373);
 *   (2) traversal order: cfg.m2l_first selects MAC-before-leaf-test (GPU kernel :504-534) or
 *       leaf-test-first (CPU :586-598);
 *   (3) cfg.tree_steps > 1 reuses the partition between rebuilds like the GPU path (:1619,1755);
 *       tree_steps = 1 is the CPU path (rebuild every call).
 * Geometry (boxes, centres, MAC) follows the reference's host arithmetic operation by operation
 * (no FMA contraction: compile with -ffp-contract=off) so that tree arrays and interaction lists
 * are BIT-EXACT against oracle/_ref on tie-free inputs; expansions are restated from the maths
 * and agree to rounding (checked in tests/test_oracle_vs_ref.py).
 *
 * Parity pinning: the reference ships no golden vectors.  This file is pinned
 * against the reference itself, compiled unmodified into oracle/_ref/libnbco_ref.so, and against
 * fixtures generated from it (tests/golden/, generator tools/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_ORDER 12

typedef struct { int x, y; } pair_t;

typedef struct orc_ctx
{
	/* configuration (reference constants.cuh:36-52) */
	int order; float radius, eps2, dens_inhom; int max_level, tree_steps, coll, unsort, m2l_first;
	/* state */
	int n, L, ntot, offM, offL, counter, rebuilt;
	float *center, *lbound, *rbound, *mpole, *local;
	int *mult, *index, *splitdim, *perm;
	pair_t *p2p, *m2l; long long p2p_n, m2l_n, p2p_cap, m2l_cap;
	pair_t *stack; long long stack_cap;
} orc_ctx;

/* ---------------- tables (mymath.cuh:26-178) ---------------- */
static double fact(int n) { double r = 1; for (int i = 2; i <= n; ++i) r *= i; return r; }
static double odfact(int n) { double r = 1; for (int i = n; i > 1; i -= 2) r *= i; return r; } /* n!!, (-1)!! = 1 */
static double binom(int n, int k) { return (k < 0 || k > n) ? 0 : fact(n) / (fact(k) * fact(n - k)); }
static double trinom(int n, int kx, int kz) { return fact(n) / (fact(kx) * fact(n - kx - kz) * fact(kz)); }
static float coeff13(int n, int m) { return (float)(((m & 1) ? -1.0 : 1.0) * odfact(2 * (n - m) - 1)); } /* fmm_cart_base3.cuh:28-33 */
static float coeff2(int n, int k) { return (float)(fact(n) / (ldexp(1.0, k) * fact(k) * fact(n - 2 * k))); } /* fmm_cart_base.cuh:33-40 */
static float ipow(float b, int e) { float r = 1.f; for (int i = 0; i < e; ++i) r *= b; return r; }

/* ---------------- tensor index maps (fmm_cart_base3.cuh:170-232) ---------------- */
static int sym_elems(int n) { return (n + 1) * (n + 2) / 2; }
static int sym_off(int p) { return p * (p + 1) * (p + 2) / 6; }
static int trl_off(int p) { return p * p; }
static int sym_idx(int x, int z, int n) { return (n * (n + 1) - (n - z) * (n - z + 1)) / 2 + n - x; }

int orc_sym_off(int p) { return sym_off(p); }
int orc_trl_off(int p) { return trl_off(p); }

/* complete a symmetric-layout tensor whose z in {0,1} entries are set, using the trace relation
 * A[x,y,z] = -A[x+2,y,z-2] - A[x,y+2,z-2] (fmm_cart_base3.cuh:611-623) */
static void trl_refine(float *A, int n)
{
	for (int z = 2; z <= n; ++z)
		for (int x = n - z; x >= 0; --x)
			A[sym_idx(x, z, n)] = -A[sym_idx(x + 2, z - 2, n)] - A[sym_idx(x, z - 2, n)];
}

/* nabla^n (1/r) times c, z in {0,1} entries from the closed form, rest by the trace relation
 * (fmm_cart_base3.cuh:698-729, 768-804).  d = unit-ish vector, r = softened distance. */
static void grad_inv_r(float *g, int n, const float d[3], float r, float c)
{
	if (n == 0) { g[0] = c / r; return; }
	float C = ((n & 1) ? -1.f : 1.f) * (1.f / ipow(r, n + 1)) * c;
	for (int z = 0; z <= 1; ++z)
		for (int x = n - z; x >= 0; --x)
		{
			int y = n - x - z;
			float t1 = 0.f;
			for (int k1 = 0; k1 <= x / 2; ++k1)
			{
				float t2 = 0.f;
				for (int k2 = 0; k2 <= y / 2; ++k2)
					t2 += coeff13(n, k1 + k2) * coeff2(y, k2) * ipow(d[1], y - 2 * k2);
				t1 += t2 * coeff2(x, k1) * ipow(d[0], x - 2 * k1);
			}
			g[sym_idx(x, z, n)] = C * t1 * ipow(d[2], z);
		}
	trl_refine(g, n);
}

/* d^(x) d^(y) d^(z) for all multi-indices of order n, symmetric layout (fmm_cart_base3.cuh:806-819) */
static void tensor_pow(float *pw, int n, const float d[3])
{
	for (int z = 0; z <= n; ++z)
		for (int x = n - z; x >= 0; --x)
			pw[sym_idx(x, z, n)] = ipow(d[0], x) * ipow(d[1], n - x - z) * ipow(d[2], z);
}

/* C (order nA-nB, traceless storage: only z in {0,1}) += c * <A (order nA, full symmetric layout),
 * B (order nB, full symmetric layout)> with trinomial weights (fmm_cart_base3.cuh:378-426) */
static void contract_trl_ma(float *Cout, const float *A, const float *B, float c, int nA, int nB)
{
	int nC = nA - nB, i = 0;
	for (int z = 0; z <= (nC < 1 ? nC : 1); ++z)
		for (int x = nC - z; x >= 0; --x)
		{
			int y = nC - x - z;
			float t = 0.f;
			for (int kz = 0; kz <= nB; ++kz)
				for (int kx = 0; kx <= nB - kz; ++kx)
				{
					int ky = nB - kx - kz;
					(void)ky;
					t += (float)trinom(nB, kx, kz) * A[sym_idx(x + kx, z + kz, nA)] * B[sym_idx(kx, kz, nB)];
				}
			(void)y;
			Cout[i++] += c * t;
		}
}

/* ---------------- operators ---------------- */

/* P2M, order q: M_q[x,y,z] += (-1)^q/q! dx^x dy^y dz^z (fmm_cart_base3.cuh:951-961) */
static void p2m_acc(float *M, int q, const float d[3])
{
	float C = (float)(((q & 1) ? -1.0 : 1.0)) * (float)(1.0 / fact(q));
	int i = 0;
	for (int z = 0; z <= q; ++z)
		for (int x = q - z; x >= 0; --x)
			M[i++] += C * ipow(d[0], x) * ipow(d[1], q - x - z) * ipow(d[2], z);
}

/* M2M, order n of the shifted tuple (fmm_cart_base3.cuh:1111-1146); d = new - old centre */
static void m2m_acc(float *Mout, const float *Mtuple, int n, const float d[3])
{
	int i = 0;
	float C = (float)(1.0 / fact(n));
	for (int z = 0; z <= n; ++z)
		for (int x = n - z; x >= 0; --x)
		{
			int y = n - x - z;
			float t = 0.f;
			for (int m = 0; m <= n; ++m)
			{
				const float *Mo = Mtuple + sym_off(n - m);
				float c = 0.f;
				for (int k1 = 0; k1 <= (x < m ? x : m); ++k1)
				{
					float c2 = 0.f;
					int lo = m - k1 - y; if (lo < 0) lo = 0;
					int hi = z < m - k1 ? z : m - k1;
					for (int k3 = lo; k3 <= hi; ++k3)
					{
						int k2 = m - k1 - k3;
						c2 += (float)binom(y, k2) * (float)binom(z, k3) * ipow(d[1], k2) * ipow(d[2], k3)
						      * Mo[sym_idx(x - k1, z - k3, n - m)];
					}
					c += c2 * (float)binom(x, k1) * ipow(d[0], k1);
				}
				t += c * (float)fact(n - m);
			}
			Mout[i++] += C * t;
		}
}

/* M2L: symmetric multipoles (orders 0..p-1, dipole skipped) -> traceless locals orders 1..p, all
 * terms with n + k <= p (static_m2l_acc3<1,-2,false,*,true>, fmm_cart_base3.cuh:1265-1346) */
static void m2l_acc(float *Ltuple, const float *Mtuple, int p, const float d[3], float r)
{
	float g[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 2) / 2];
	for (int m = 1; m <= p; ++m)
	{
		float scal = ipow(r, m + 1); /* rescaling against fp32 overflow (:1288) */
		grad_inv_r(g, m, d, r, scal);
		for (int n = 1; n <= m; ++n)
		{
			int mn = m - n;
			if (mn == 1) continue; /* dipole is identically zero about the centre of charge */
			float C = (float)(1.0 / fact(n)) / scal;
			contract_trl_ma(Ltuple + trl_off(n), g, Mtuple + sym_off(mn), C, m, mn);
		}
	}
}

/* expand the traceless-stored orders 1..p of a local tuple into full symmetric layout */
static void local_expand(float *S, const float *Ltrl, int p)
{
	for (int q = 1; q <= p; ++q)
	{
		float *dst = S + sym_off(q);
		for (int j = 0; j < 2 * q + 1; ++j)
			dst[j] = Ltrl[trl_off(q) + j];
		trl_refine(dst, q);
	}
}

/* L2L: L'_n += sum_{m=n..p} C(m,m-n) <L_m, d^(m-n)>, n = 1..p (fmm_cart_base3.cuh:1383-1412,
 * symmetric-power variant used by fmm_pushl3_kdtree_krnl, fmm_cart3_kdtree.cuh:1134-1169) */
static void l2l_acc(float *Lout_trl, const float *Lsym, int p, const float d[3])
{
	float pw[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 2) / 2];
	for (int n = 1; n <= p; ++n)
		for (int m = n; m <= p; ++m)
		{
			tensor_pow(pw, m - n, d);
			contract_trl_ma(Lout_trl + trl_off(n), Lsym + sym_off(m), pw, (float)binom(m, m - n), m, m - n);
		}
}

/* L2P: field = -sum_n n <L_n, d^(n-1)> (fmm_cart_base3.cuh:1550-1578) */
static void l2p_field(float out[3], const float *Lsym, int p, const float d[3])
{
	float pw[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 2) / 2];
	float t[3] = {0.f, 0.f, 0.f};
	for (int n = 1; n <= p; ++n)
	{
		tensor_pow(pw, n - 1, d);
		contract_trl_ma(t, Lsym + sym_off(n), pw, (float)n, n, n - 1);
	}
	out[0] = -t[0]; out[1] = -t[1]; out[2] = -t[2];
}

/* ---------------- kd-tree (fmm_cart3_kdtree.cuh:33-202) ---------------- */
static int kd_beg(int l) { return (1 << l) - 1; }
static int kd_cnt(int l) { return 1 << l; }

int orc_kd_levels(int n, int order, float dens_inhom, int max_level)
/* fmm_cart3_kdtree.cuh:1507-1516 */
{
	float s = (float)(order * order);
	int L;
	if (max_level == 0)
		L = (int)roundf(log2f(dens_inhom * (float)n / s));
	else
		L = max_level;
	if (L < 2) L = 2;
	if (L > 30) L = 30;
	while (kd_cnt(L) > n) --L;
	return L;
}

static int seg_start(long long n, int i, int m) { return i == 0 ? 0 : (int)((n * i - 1) / m + 1); } /* :117-118 */

static uint32_t ordered_bits(float f)
/* monotone map float -> uint32 (:175-185); -0.0 sorts before +0.0 */
{
	uint32_t u; memcpy(&u, &f, 4);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

static int widest_axis(const float *lb, const float *rb)
{
	float dx = rb[0] - lb[0], dy = rb[1] - lb[1], dz = rb[2] - lb[2];
	return (dx > dy) ? ((dx > dz) ? 0 : 2) : ((dy > dz) ? 1 : 2); /* :92,129 */
}

typedef struct { uint32_t key; int idx; } kv_t;

static void stable_sort_kv(kv_t *a, kv_t *tmp, int n)
{
	/* bottom-up merge sort, stable */
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
		memcpy(a, tmp, sizeof(kv_t) * (size_t)n);
	}
}

static void eval_box_level(orc_ctx *c, const float *p, int l)
/* evalBox_krnl (:109-137): boxes, split axes and start indices of level l from level l-1 */
{
	int m = kd_cnt(l), beg = kd_beg(l), n = c->n;
	for (int i = 0; i < m; ++i)
	{
		int start = seg_start(n, i, m), end = seg_start(n, i + 1, m);
		int j = beg + i, parent = (j - 1) >> 1, split = c->splitdim[parent];
		float lb[3], rb[3];
		memcpy(lb, c->lbound + 3 * parent, 12); memcpy(rb, c->rbound + 3 * parent, 12);
		if (j == 2 * parent + 2) lb[split] = p[3 * (size_t)start + split];
		if (j == 2 * parent + 1) rb[split] = p[3 * (size_t)(end - 1) + split];
		memcpy(c->lbound + 3 * j, lb, 12); memcpy(c->rbound + 3 * j, rb, 12);
		c->splitdim[j] = widest_axis(lb, rb);
		c->index[j] = start;
	}
}

static void build_tree(orc_ctx *c, float *p, float *vel_or_null)
/* :1857-1873; p is permuted into tree order, c->perm[sorted] = input position */
{
	int n = c->n, L = c->L;
	kv_t *kv = malloc(sizeof(kv_t) * (size_t)n), *tmp = malloc(sizeof(kv_t) * (size_t)n);
	float *pt = malloc(12 * (size_t)n);
	int *it = malloc(sizeof(int) * (size_t)n);
	float mn[3] = {p[0], p[1], p[2]}, mx[3] = {p[0], p[1], p[2]};
	for (int i = 1; i < n; ++i)
		for (int k = 0; k < 3; ++k)
		{
			mn[k] = fminf(mn[k], p[3 * (size_t)i + k]);
			mx[k] = fmaxf(mx[k], p[3 * (size_t)i + k]);
		}
	memcpy(c->lbound, mn, 12); memcpy(c->rbound, mx, 12);
	c->splitdim[0] = widest_axis(mn, mx);
	c->index[0] = 0;
	for (int i = 0; i < n; ++i) c->perm[i] = i;
	for (int l = 0; l <= L - 1; ++l)
	{
		if (l > 0) eval_box_level(c, p, l);
		int m = kd_cnt(l);
		for (int s = 0; s < m; ++s)
		{
			int a = seg_start(n, s, m), b = seg_start(n, s + 1, m), ax = c->splitdim[kd_beg(l) + s];
			for (int i = a; i < b; ++i) { kv[i].key = ordered_bits(p[3 * (size_t)i + ax]); kv[i].idx = i; }
			stable_sort_kv(kv + a, tmp + a, b - a);
		}
		for (int i = 0; i < n; ++i) { memcpy(pt + 3 * (size_t)i, p + 3 * (size_t)kv[i].idx, 12); it[i] = c->perm[kv[i].idx]; }
		memcpy(p, pt, 12 * (size_t)n); memcpy(c->perm, it, sizeof(int) * (size_t)n);
	}
	eval_box_level(c, p, L);
	/* multLeaves (appel.cuh:184-197) */
	int beg = kd_beg(L), m = kd_cnt(L);
	for (int i = 0; i < m; ++i)
		c->mult[beg + i] = (i == m - 1 ? n : c->index[beg + i + 1]) - c->index[beg + i];
	if (vel_or_null)
	{
		for (int i = 0; i < n; ++i) memcpy(pt + 3 * (size_t)i, vel_or_null + 3 * (size_t)c->perm[i], 12);
		memcpy(vel_or_null, pt, 12 * (size_t)n);
	}
	free(kv); free(tmp); free(pt); free(it);
}

static float kd_size(const float *l, const float *r)
{
	float dx = r[0] - l[0], dy = r[1] - l[1], dz = r[2] - l[2];
	return dx * dx + dy * dy + dz * dz; /* helper_math.h dot(): (x*x + y*y) + z*z, no contraction */
}

static int kd_admissible(const orc_ctx *c, int n1, int n2)
/* :401-414, host branch (pow in SCAL = powf) */
{
	const float *c1 = c->center + 3 * n1, *c2 = c->center + 3 * n2;
	float dx = c2[0] - c1[0], dy = c2[1] - c1[1], dz = c2[2] - c1[2];
	float dist2 = dx * dx + dy * dy + dz * dz;
	float sz1 = kd_size(c->lbound + 3 * n1, c->rbound + 3 * n1), sz2 = kd_size(c->lbound + 3 * n2, c->rbound + 3 * n2);
	int mm = c->mult[n1] > c->mult[n2] ? c->mult[n1] : c->mult[n2];
	float M = powf((float)mm / c->mult[0], 1.f / (3 * c->order + 6));
	float parM = c->radius * M;
	return parM * parM * fmaxf(sz1, sz2) < dist2;
}

static int push_pair(pair_t **list, long long *n, long long *cap, int x, int y)
{
	if (*n == *cap)
	{
		*cap = *cap ? *cap * 2 : 1024;
		*list = realloc(*list, sizeof(pair_t) * (size_t)*cap);
		if (!*list) return -1;
	}
	(*list)[*n].x = x; (*list)[*n].y = y; ++*n;
	return 0;
}

static void dual_traversal(orc_ctx *c)
/* fmm_dualTraversal_cpu (:569-611); m2l_first swaps the first and third tests like the GPU
 * kernel instantiation fmm_dualTraversal<true> (:504-534) */
{
	int ntot = c->ntot;
	long long top = 0;
	c->p2p_n = c->m2l_n = 0;
#define PUSH(a, b) push_pair(&c->stack, &top, &c->stack_cap, a, b)
	PUSH(0, 0);
	while (top > 0)
	{
		pair_t np = c->stack[--top];
		int xl = 2 * np.x + 1 >= ntot, yl = 2 * np.y + 1 >= ntot;
		if (!c->m2l_first && xl && yl)
		{
			if (np.x != np.y) push_pair(&c->p2p, &c->p2p_n, &c->p2p_cap, np.x, np.y);
		}
		else if (np.x == np.y && !xl)
		{
			PUSH(2 * np.x + 1, 2 * np.x + 1);
			PUSH(2 * np.x + 1, 2 * np.x + 2);
			PUSH(2 * np.x + 2, 2 * np.x + 2);
		}
		else if (np.x != np.y && kd_admissible(c, np.x, np.y))
			push_pair(&c->m2l, &c->m2l_n, &c->m2l_cap, np.x, np.y);
		else if (xl && yl)
		{
			if (np.x != np.y) push_pair(&c->p2p, &c->p2p_n, &c->p2p_cap, np.x, np.y);
		}
		else if (xl || (!yl && kd_size(c->lbound + 3 * np.x, c->rbound + 3 * np.x) <= kd_size(c->lbound + 3 * np.y, c->rbound + 3 * np.y)))
		{
			PUSH(np.x, 2 * np.y + 1);
			PUSH(np.x, 2 * np.y + 2);
		}
		else
		{
			PUSH(2 * np.x + 1, np.y);
			PUSH(2 * np.x + 2, np.y);
		}
	}
#undef PUSH
}

/* ---------------- near field (fmm_cart3_kdtree.cuh:767-795, direct.cuh:27-31) ---------------- */
static void p2p_one_way(float *a1, const float *p1, const float *p2, int m1, int m2, float eps2)
{
	for (int h = 0; h < m1; ++h)
	{
		float ax = 0.f, ay = 0.f, az = 0.f;
		for (int g = 0; g < m2; ++g)
		{
			float dx = p1[3 * h] - p2[3 * g], dy = p1[3 * h + 1] - p2[3 * g + 1], dz = p1[3 * h + 2] - p2[3 * g + 2];
			float dist2 = dx * dx + dy * dy + dz * dz + eps2;
			float inv2 = 1.f / dist2;
			double inv = sqrt((double)inv2);
			float k = (float)((double)inv2 * inv);
			ax = fmaf(k, dx, ax); ay = fmaf(k, dy, ay); az = fmaf(k, dz, az);
		}
		a1[3 * h] += ax; a1[3 * h + 1] += ay; a1[3 * h + 2] += az;
	}
}

/* ---------------- context ---------------- */

orc_ctx *orc_create(int order, float radius, float eps2, float dens_inhom, int max_level, int tree_steps,
                    int coll, int unsort, int m2l_first)
{
	if (order < 1 || order > ORC_MAX_ORDER) return NULL;
	orc_ctx *c = calloc(1, sizeof(orc_ctx));
	c->order = order; c->radius = radius; c->eps2 = eps2; c->dens_inhom = dens_inhom; c->max_level = max_level;
	c->tree_steps = tree_steps < 1 ? 1 : tree_steps; c->coll = coll; c->unsort = unsort; c->m2l_first = m2l_first;
	return c;
}

static void free_tree(orc_ctx *c)
{
	free(c->center); free(c->lbound); free(c->rbound); free(c->mpole); free(c->local);
	free(c->mult); free(c->index); free(c->splitdim); free(c->perm);
	c->center = c->lbound = c->rbound = c->mpole = c->local = NULL;
	c->mult = c->index = c->splitdim = c->perm = NULL;
}

void orc_destroy(orc_ctx *c)
{
	if (!c) return;
	free_tree(c); free(c->p2p); free(c->m2l); free(c->stack); free(c);
}

static void plan(orc_ctx *c, int n)
{
	int L = orc_kd_levels(n, c->order, c->dens_inhom, c->max_level);
	if (n == c->n && L == c->L && c->center) return;
	free_tree(c);
	c->n = n; c->L = L; c->ntot = (1 << (L + 1)) - 1; c->counter = 0;
	c->offM = sym_off(c->order); c->offL = trl_off(c->order + 1);
	size_t nt = (size_t)c->ntot;
	c->center = calloc(nt, 12); c->lbound = calloc(nt, 12); c->rbound = calloc(nt, 12);
	c->mpole = calloc(nt * c->offM, 4); c->local = calloc(nt * c->offL, 4);
	c->mult = calloc(nt + 1, 4); c->index = calloc(nt + 1, 4); c->splitdim = calloc(nt, 4);
	c->perm = calloc((size_t)n, 4);
}

/* fmm_cart3_kdtree[_cpu]: pos (n float3; followed by n float3 velocities when vel != NULL and
 * unsort == 0), acc out.  param may be NULL. */
int orc_fmm3_kd(orc_ctx *c, float *pos, float *vel, float *acc, int n, const float *param)
{
	plan(c, n);
	int p = c->order, L = c->L, ntot = c->ntot, offM = c->offM, offL = c->offL;
	int rebuild = c->unsort || (c->counter % c->tree_steps == 0);
	c->rebuilt = rebuild;
	if (rebuild)
		build_tree(c, pos, (!c->unsort) ? vel : NULL);
	int beg = kd_beg(L), m = kd_cnt(L);

	/* centerLeaves (appel.cuh:226-243) */
	for (int i = beg; i < beg + m; ++i)
	{
		float t[3] = {0.f, 0.f, 0.f};
		const float *pi = pos + 3 * (size_t)c->index[i];
		for (int j = 0; j < c->mult[i]; ++j) { t[0] += pi[3 * j]; t[1] += pi[3 * j + 1]; t[2] += pi[3 * j + 2]; }
		if (c->mult[i] > 0) { t[0] /= (float)c->mult[i]; t[1] /= (float)c->mult[i]; t[2] /= (float)c->mult[i]; }
		memcpy(c->center + 3 * i, t, 12);
	}
	memset(c->mpole, 0, sizeof(float) * (size_t)ntot * offM);
	memset(c->local, 0, sizeof(float) * (size_t)ntot * offL);
	/* P2M (:231-250) */
	for (int i = beg; i < beg + m; ++i)
	{
		float *M = c->mpole + (size_t)i * offM;
		const float *pi = pos + 3 * (size_t)c->index[i];
		M[0] = (float)c->mult[i];
		if (p >= 3)
			for (int j = 0; j < c->mult[i]; ++j)
			{
				float d[3] = {pi[3 * j] - c->center[3 * i], pi[3 * j + 1] - c->center[3 * i + 1], pi[3 * j + 2] - c->center[3 * i + 2]};
				for (int q = 2; q <= p - 1; ++q) p2m_acc(M + sym_off(q), q, d);
			}
	}
	/* M2M (:328-368), level L-1 -> 0 */
	for (int l = L - 1; l >= 0; --l)
		for (int i = kd_beg(l); i < kd_beg(l + 1); ++i)
		{
			int ch[2] = {2 * i + 1, 2 * i + 2};
			int mlt = c->mult[ch[0]] + c->mult[ch[1]];
			float m0 = (float)mlt, co[3] = {0.f, 0.f, 0.f};
			for (int k = 0; k < 2; ++k)
				for (int a = 0; a < 3; ++a)
					co[a] += (float)c->mult[ch[k]] * c->center[3 * ch[k] + a];
			for (int a = 0; a < 3; ++a) co[a] /= m0;
			float *M = c->mpole + (size_t)i * offM;
			if (p >= 3)
				for (int k = 0; k < 2; ++k)
				{
					float d[3] = {co[0] - c->center[3 * ch[k]], co[1] - c->center[3 * ch[k] + 1], co[2] - c->center[3 * ch[k] + 2]};
					for (int q = 2; q <= p - 1; ++q)
						m2m_acc(M + sym_off(q), c->mpole + (size_t)ch[k] * offM, q, d);
				}
			M[0] = m0;
			memcpy(c->center + 3 * i, co, 12);
			c->mult[i] = mlt;
		}

	dual_traversal(c);

	memset(acc, 0, 12 * (size_t)n);
	if (c->coll)
	{
		for (long long k = 0; k < c->p2p_n; ++k)
		{
			int n1 = c->p2p[k].x, n2 = c->p2p[k].y;
			int i1 = c->index[n1], i2 = c->index[n2], m1 = c->mult[n1], m2 = c->mult[n2];
			p2p_one_way(acc + 3 * (size_t)i1, pos + 3 * (size_t)i1, pos + 3 * (size_t)i2, m1, m2, c->eps2);
			p2p_one_way(acc + 3 * (size_t)i2, pos + 3 * (size_t)i2, pos + 3 * (size_t)i1, m2, m1, c->eps2);
		}
		for (int i = beg; i < beg + m; ++i)
		{
			int i1 = c->index[i], m1 = c->mult[i];
			p2p_one_way(acc + 3 * (size_t)i1, pos + 3 * (size_t)i1, pos + 3 * (size_t)i1, m1, m1, c->eps2);
		}
	}
	/* M2L (:613-671, host branch :667-668) */
	for (long long k = 0; k < c->m2l_n; ++k)
	{
		int n1 = c->m2l[k].x, n2 = c->m2l[k].y;
		float d[3] = {c->center[3 * n1] - c->center[3 * n2], c->center[3 * n1 + 1] - c->center[3 * n2 + 1], c->center[3 * n1 + 2] - c->center[3 * n2 + 2]};
		float r = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + c->eps2);
		d[0] /= r; d[1] /= r; d[2] /= r;
		float nd[3] = {-d[0], -d[1], -d[2]};
		m2l_acc(c->local + (size_t)n1 * offL, c->mpole + (size_t)n2 * offM, p, d, r);
		m2l_acc(c->local + (size_t)n2 * offL, c->mpole + (size_t)n1 * offM, p, nd, r);
	}
	/* L2L (:1713-1730), level 1 -> L-1 pushes to its children */
	float *S = malloc(sizeof(float) * (size_t)sym_off(p + 1));
	for (int l = 1; l <= L - 1; ++l)
		for (int i = kd_beg(l); i < kd_beg(l + 1); ++i)
		{
			local_expand(S, c->local + (size_t)i * offL, p);
			for (int k = 1; k <= 2; ++k)
			{
				int ch = 2 * i + k;
				float d[3] = {c->center[3 * ch] - c->center[3 * i], c->center[3 * ch + 1] - c->center[3 * i + 1], c->center[3 * ch + 2] - c->center[3 * i + 2]};
				l2l_acc(c->local + (size_t)ch * offL, S, p, d);
			}
		}
	/* L2P (:1227-1253) */
	for (int i = beg; i < beg + m; ++i)
	{
		local_expand(S, c->local + (size_t)i * offL, p);
		const float *pi = pos + 3 * (size_t)c->index[i];
		float *ai = acc + 3 * (size_t)c->index[i];
		for (int j = 0; j < c->mult[i]; ++j)
		{
			float d[3] = {pi[3 * j] - c->center[3 * i], pi[3 * j + 1] - c->center[3 * i + 1], pi[3 * j + 2] - c->center[3 * i + 2]};
			float f[3];
			l2p_field(f, S, p, d);
			ai[3 * j] += f[0]; ai[3 * j + 1] += f[1]; ai[3 * j + 2] += f[2];
		}
	}
	free(S);
	if (param) /* rescale (appel.cuh:506-512) */
		for (size_t i = 0; i < 3 * (size_t)n; ++i) acc[i] *= param[0];
	if (c->unsort)
	{
		/* gather_inverse (:1913-1917): out[perm[i]] = in[i] */
		float *t = malloc(12 * (size_t)n);
		for (int i = 0; i < n; ++i) memcpy(t + 3 * (size_t)c->perm[i], pos + 3 * (size_t)i, 12);
		memcpy(pos, t, 12 * (size_t)n);
		for (int i = 0; i < n; ++i) memcpy(t + 3 * (size_t)c->perm[i], acc + 3 * (size_t)i, 12);
		memcpy(acc, t, 12 * (size_t)n);
		free(t);
	}
	++c->counter;
	return 0;
}

/* ---------------- direct sum, elastic term, step, integrators ---------------- */

void orc_direct3(const float *p, float *a, int n, const float *param, float eps2)
/* direct3_core (direct.cuh:192-226): Kahan-compensated, IEEE divide and sqrt */
{
	float k = param ? param[0] : 1.f;
	for (int i = 0; i < n; ++i)
	{
		float s[3] = {0.f, 0.f, 0.f}, c[3] = {0.f, 0.f, 0.f};
		for (int j = 0; j < n; ++j)
		{
			float d[3] = {p[3 * i] - p[3 * j], p[3 * i + 1] - p[3 * j + 1], p[3 * i + 2] - p[3 * j + 2]};
			float dist2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + eps2;
			float inv2 = 1.f / dist2, w = sqrtf(inv2);
			for (int q = 0; q < 3; ++q)
			{
				float y = d[q] * inv2 * w - c[q];
				float t = s[q] + y;
				c[q] = (t - s[q]) - y;
				s[q] = t;
			}
		}
		a[3 * i] = k * s[0]; a[3 * i + 1] = k * s[1]; a[3 * i + 2] = k * s[2];
	}
}

void orc_add_elastic(const float *p, float *a, int n, const float *k3)
/* add_elastic_cpu (kernel.cuh:154-173) */
{
	for (int i = 0; i < n; ++i)
		for (int q = 0; q < 3; ++q)
			a[3 * i + q] -= p[3 * i + q] * (k3 ? k3[q] : 1.f);
}

void orc_step(float *b, const float *a, float ds, int n)
/* step_cpu (kernel.cuh:106-117) */
{
	for (size_t i = 0; i < 3 * (size_t)n; ++i) b[i] += a[i] * ds;
}

/* evaluator: 0 direct3, 1 fmm3_kd, 2 direct3 + elastic, 3 fmm3_kd + elastic (main3.cu:47-69) */
int orc_eval(orc_ctx *c, int evaluator, float *buf, int n, const float *param)
{
	float *pos = buf, *vel = buf + 3 * (size_t)n, *acc = buf + 6 * (size_t)n;
	switch (evaluator)
	{
		case 0: orc_direct3(pos, acc, n, param, c->eps2); return 0;
		case 1: return orc_fmm3_kd(c, pos, vel, acc, n, param);
		case 2: orc_direct3(pos, acc, n, param, c->eps2); orc_add_elastic(pos, acc, n, param ? param + 3 : NULL); return 0;
		case 3: { int s = orc_fmm3_kd(c, pos, vel, acc, n, param); orc_add_elastic(pos, acc, n, param ? param + 3 : NULL); return s; }
		default: return -1;
	}
}

int orc_integrate(orc_ctx *c, int scheme, int evaluator, float *buf, int n, const float *param, double dt_, long long nsteps)
/* integrator.cuh:32-167; coefficients formed in long double and cast to float at the call */
{
	float *pos = buf, *vel = buf + 3 * (size_t)n, *acc = buf + 6 * (size_t)n;
	const long double dt = (float)dt_;
#define K(cf) orc_step(vel, acc, (float)(cf), n)
#define D(cf) orc_step(pos, vel, (float)(cf), n)
#define F() do { int s_ = orc_eval(c, evaluator, buf, n, param); if (s_) return s_; } while (0)
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

/* energy, double accumulation: out = {kinetic, elastic, param[0] * pair} */
void orc_energy(const float *buf, int n, const float *param, float eps2, double *out)
{
	const float *x = buf, *v = buf + 3 * (size_t)n;
	double ke = 0, el = 0, pe = 0;
	for (int i = 0; i < n; ++i)
		for (int q = 0; q < 3; ++q)
		{
			ke += 0.5 * (double)v[3 * i + q] * v[3 * i + q];
			el += 0.5 * (double)(param ? param[3 + q] : 1.f) * x[3 * i + q] * x[3 * i + q];
		}
	for (int i = 0; i < n; ++i)
		for (int j = i + 1; j < n; ++j)
		{
			double dx = (double)x[3 * i] - x[3 * j], dy = (double)x[3 * i + 1] - x[3 * j + 1], dz = (double)x[3 * i + 2] - x[3 * j + 2];
			pe += 1.0 / sqrt(dx * dx + dy * dy + dz * dz + (double)eps2);
		}
	out[0] = ke; out[1] = el; out[2] = (param ? (double)param[0] : 1.0) * pe;
}

double orc_mean_rel_err(const float *x, const float *ref, int n, double *max_out)
/* rel_diff1 (reductions.cuh:37-42), mean and max */
{
	double s = 0, mx = 0;
	for (int i = 0; i < n; ++i)
	{
		float dx = x[3 * i] - ref[3 * i], dy = x[3 * i + 1] - ref[3 * i + 1], dz = x[3 * i + 2] - ref[3 * i + 2];
		float d2 = dx * dx + dy * dy + dz * dz;
		float r2 = ref[3 * i] * ref[3 * i] + ref[3 * i + 1] * ref[3 * i + 1] + ref[3 * i + 2] * ref[3 * i + 2] + 1.e-18f;
		double e = sqrt(fmax((double)(d2 / r2), 0.0));
		s += e; if (e > mx) mx = e;
	}
	if (max_out) *max_out = mx;
	return n > 0 ? s / n : 0.0;
}

/* ---------------- introspection ---------------- */
void orc_info(const orc_ctx *c, long long *out)
/* out[0..7] = L, order, n, nodes, p2p pairs, m2l pairs, offM, offL; out[8] = rebuilt */
{
	out[0] = c->L; out[1] = c->order; out[2] = c->n; out[3] = c->ntot; out[4] = c->p2p_n; out[5] = c->m2l_n;
	out[6] = c->offM; out[7] = c->offL; out[8] = c->rebuilt;
}

void orc_get_tree(const orc_ctx *c, float *center, float *lbound, float *rbound, float *mpole, float *local,
                  int *mult, int *index, int *splitdim, int *perm)
{
	size_t nt = (size_t)c->ntot;
	if (center) memcpy(center, c->center, 12 * nt);
	if (lbound) memcpy(lbound, c->lbound, 12 * nt);
	if (rbound) memcpy(rbound, c->rbound, 12 * nt);
	if (mpole) memcpy(mpole, c->mpole, 4 * nt * c->offM);
	if (local) memcpy(local, c->local, 4 * nt * c->offL);
	if (mult) memcpy(mult, c->mult, 4 * nt);
	if (index) memcpy(index, c->index, 4 * nt);
	if (splitdim) memcpy(splitdim, c->splitdim, 4 * nt);
	if (perm) memcpy(perm, c->perm, 4 * (size_t)c->n);
}

static int pair_cmp(const void *a, const void *b)
{
	const pair_t *p = a, *q = b;
	if (p->x != q->x) return p->x < q->x ? -1 : 1;
	return (p->y > q->y) - (p->y < q->y);
}

void orc_get_lists(const orc_ctx *c, int *p2p, int *m2l)
/* sorted ascending by (x, y) so that lists compare as sets */
{
	if (p2p) { memcpy(p2p, c->p2p, sizeof(pair_t) * (size_t)c->p2p_n); qsort(p2p, (size_t)c->p2p_n, sizeof(pair_t), pair_cmp); }
	if (m2l) { memcpy(m2l, c->m2l, sizeof(pair_t) * (size_t)c->m2l_n); qsort(m2l, (size_t)c->m2l_n, sizeof(pair_t), pair_cmp); }
}

/* ---------------- operator-level entry points (unit tests of the CUDA templates) ---------------- */
void orc_op_p2m(float *M, int p, const float *d) { for (int q = 2; q <= p - 1; ++q) p2m_acc(M + sym_off(q), q, d); }
void orc_op_m2m(float *Mout, const float *Min, int p, const float *d) { for (int q = 2; q <= p - 1; ++q) m2m_acc(Mout + sym_off(q), Min, q, d); }
void orc_op_m2l(float *L, const float *M, int p, const float *d_unit, float r) { m2l_acc(L, M, p, d_unit, r); }
void orc_op_l2l(float *Lchild, const float *Lparent, int p, const float *d)
{
	float S[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 2) * (ORC_MAX_ORDER + 3) / 6];
	local_expand(S, Lparent, p);
	l2l_acc(Lchild, S, p, d);
}
void orc_op_l2p(float *f, const float *L, int p, const float *d)
{
	float S[(ORC_MAX_ORDER + 1) * (ORC_MAX_ORDER + 2) * (ORC_MAX_ORDER + 3) / 6];
	local_expand(S, L, p);
	l2p_field(f, S, p, d);
}
