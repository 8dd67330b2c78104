// oracle/ref_harness.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" shell around the UNMODIFIED reference sources, compiled where they lie
// (/root/reference/Simulation) into oracle/_ref/libnbco_ref.so by oracle/Makefile.
// It is used to (1) pin oracle/nbco_oracle.c (our own CPU restatement) and (2) serve as
// the "reference" CPU baseline in bench.py.  Nothing here re-implements reference maths:
// every number comes out of a reference function.  The only logic of our own is
//   * the phase driver ref_fmm3_phases(), which calls the reference's CPU phase functions
//     in the order of fmm_cart3_kdtree.cuh:1857-1908 so that intermediates can be dumped;
//   * a dual traversal with the MAC test before the leaf test (the order of the reference
//     GPU kernel, fmm_cart3_kdtree.cuh:504-534), built on the reference's own
//     kd_admissible()/kd_size(), because the CPU reference only ships the other order.
//
// The reference TU is main3.cu (it owns initGA/initU/coulombOscillator*): its main() is
// renamed so that it can be called as a function.

#define main nbco_ref_cli_main
#include "main3.cu"
#undef main

#include <cstring>

namespace {

typedef void (*eval_fn)(VEC*, VEC*, int, const SCAL*);

eval_fn pick_eval(int which)
{
	switch (which)
	{
		case 0: return direct3_cpu;                  // direct.cuh:247
		case 1: return fmm_cart3_kdtree_cpu;         // fmm_cart3_kdtree.cuh:1773
		case 2: return coulombOscillatorDirect_cpu;  // main3.cu:53
		case 3: return coulombOscillatorFMMKD3_cpu;  // main3.cu:65
		default: return nullptr;
	}
}

} // namespace

extern "C" {

// ---- configuration: writes the reference's mutable globals (constants.cuh:36-52) ----
void ref_config(int order, float radius, float eps2, float dens, int threads, int coll_, int unsort, int tsteps)
{
	::fmm_order = order;
	::tree_radius = radius;
	::EPS2 = eps2;
	::dens_inhom = dens;
	::CPU_THREADS = threads;
	::coll = coll_ != 0;
	::b_unsort = unsort != 0;
	::tree_steps = tsteps;
}

int ref_cli(int argc, const char **argv) { return nbco_ref_cli_main(argc, argv); }

// ---- initial conditions, exactly as main3.cu:662-666 ----
void ref_init_ga(float *buf, int n, const float *x3, const float *u3)
{
	std::mt19937_64 gen(5351550349027530206ULL);
	gen.discard(624*2);
	initGA((VEC*)buf, 2*n, VEC{x3[0], x3[1], x3[2]}, VEC{u3[0], u3[1], u3[2]}, gen);
}

void ref_init_test_cube(float *buf, int n, const float *x3, const float *u3)
// the "-test" initial state: initGA followed by initU on the positions (main3.cu:664-666)
{
	std::mt19937_64 gen(5351550349027530206ULL);
	gen.discard(624*2);
	initGA((VEC*)buf, 2*n, VEC{x3[0], x3[1], x3[2]}, VEC{u3[0], u3[1], u3[2]}, gen);
	initU((VEC*)buf, 2*n, VEC{-1, -1, -1}, VEC{1, 1, 1}, gen);
}

// ---- evaluators through the reference's own plugin type (integrator.cuh:22) ----
int ref_eval(int which, float *buf, int n, const float *param)
// buf = [pos | vel | acc], 3*n float3; acc is written
{
	eval_fn f = pick_eval(which);
	if (!f) return -1;
	compute_force(f, buf, n, param);
	return 0;
}

int ref_integrate(int scheme, int which, float *buf, int n, const float *param, double dt, int nsteps)
// scheme: 0 symplectic_euler, 1 leapfrog, 2 forestruth, 3 pefrl (integrator.cuh:32-167).
// The caller is responsible for the initial compute_force (main3.cu:835-839).
{
	eval_fn f = pick_eval(which);
	if (!f) return -1;
	SCAL dts = (SCAL)dt; // main3.cu:231 stores dt as SCAL before widening it to long double
	for (int s = 0; s < nsteps; ++s)
		switch (scheme)
		{
			case 0: symplectic_euler(f, buf, n, param, dts, step_cpu, 1); break;
			case 1: leapfrog(f, buf, n, param, dts, step_cpu, 1); break;
			case 2: forestruth(f, buf, n, param, dts, step_cpu, 1); break;
			case 3: pefrl(f, buf, n, param, dts, step_cpu, 1); break;
			default: return -1;
		}
	return 0;
}

void ref_step(float *b, const float *a, float ds, int n) { step_cpu((VEC*)b, (const VEC*)a, ds, n); }
void ref_add_elastic(float *p, float *a, int n, const float *k3) { add_elastic_cpu((VEC*)p, (VEC*)a, n, k3); }

double ref_mean_rel_err(const float *x, const float *ref, int n)
// mean rel_diff1 (reductions.cuh:37-42), accumulated in double by us
{
	double s = 0;
	for (int i = 0; i < n; ++i)
		s += rel_diff1(((const VEC*)x)[i], ((const VEC*)ref)[i]);
	return s / n;
}

int ref_kd_levels(int n)
// depth rule of fmm_cart3_kdtree_cpu, fmm_cart3_kdtree.cuh:1789-1796
{
	int order = ::fmm_order;
	SCAL s = order*order;
	int L = (int)std::round(std::log2(::dens_inhom*(SCAL)n/s));
	L = std::max(L, 2);
	L = std::min(L, 30);
	while (kd_n(L) > n)
		--L;
	return L;
}

int ref_sym_offset(int p) { return symmetricoffset3(p); }
int ref_trl_offset(int p) { return tracelessoffset3(p); }

// ---- phase-by-phase FMM with every intermediate exposed ----
// Order of calls = fmm_cart3_kdtree_cpu, fmm_cart3_kdtree.cuh:1833-1908.
// p (n float3) is permuted in place into tree order; unsort[n] receives the permutation
// (sorted position -> input position).  Tree arrays are sized by the caller from
// ref_kd_levels(): ntot = 2^(L+1)-1.  Lists are returned through caller buffers of
// capacity cap (pairs); the true counts are written to counts[0] (p2p) and counts[1] (m2l).
// m2l_first != 0 selects the GPU kernel's test order (MAC before leaf test).
// acc is left in TREE order, scaled by param[0] when param != NULL.
int ref_fmm3_phases(float *p_, float *acc_, int n, const float *param, int m2l_first,
                    int *unsort, float *center, float *lbound, float *rbound, float *mpole, float *local,
                    int *mult, int *index, int *splitdim,
                    int *p2p_pairs, int *m2l_pairs, long long cap, long long *counts)
{
	VEC *p = (VEC*)p_, *a = (VEC*)acc_;
	int order = ::fmm_order;
	int L = ref_kd_levels(n);
	int ntot = kd_ntot(L);
	fmmTree_kd tree;
	tree.center = (VEC*)center; tree.lbound = (VEC*)lbound; tree.rbound = (VEC*)rbound;
	tree.mpole = mpole; tree.local = local; tree.mult = mult; tree.index = index; tree.splitdim = splitdim;
	tree.p = order;

	std::vector<unsigned long long> keys(n);
	std::vector<int> ind(n);
	std::vector<char> c_tmp((size_t)n*sizeof(VEC));

	// bounding box (fmm_cart3_kdtree.cuh:1833-1855, single pass; min/max are exact in any order)
	VEC mm[2] = {p[0], p[0]};
	for (int j = 1; j < n; ++j)
	{
		mm[0] = fmin(mm[0], p[j]);
		mm[1] = fmax(mm[1], p[j]);
	}
	evalRootBox_cpu(tree, mm);
	evalKeys_kdtree_cpu(keys.data(), tree.splitdim, p, n, 0);
	evalIndices_cpu(ind.data(), n);
	evalIndices_cpu(unsort, n);
	sort_particle_cpu(p, c_tmp.data(), n, keys.data(), ind.data(), unsort);
	for (int l = 1; l <= L-1; ++l)
	{
		evalBox_cpu(tree, p, n, l);
		evalKeys_kdtree_cpu(keys.data(), tree.splitdim + kd_beg(l), p, n, l);
		evalIndices_cpu(ind.data(), n);
		sort_particle_cpu(p, c_tmp.data(), n, keys.data(), ind.data(), unsort);
	}
	evalBox_cpu<true>(tree, p, n, L);
	fmm_init3_kdtree_cpu(tree, L);
	int beg = kd_beg(L), m = kd_n(L);
	multLeaves_cpu(tree.mult + beg, tree.index + beg, m, n);
	centerLeaves_cpu(tree.center + beg, tree.mult + beg, tree.index + beg, p, m);
	fmm_multipoleLeaves3_kdtree_cpu(tree, p, L);
	for (int l = L-1; l >= 0; --l)
		fmm_buildTree3_kdtree_cpu(tree, l);

	std::vector<int2> p2p_list, m2l_list, stack;
	SCAL radius = ::tree_radius;
	if (!m2l_first)
		fmm_dualTraversal_cpu(tree, p2p_list, m2l_list, stack, radius, L);
	else
	{
		// same walk as fmm_dualTraversal_cpu (fmm_cart3_kdtree.cuh:569-611) with the tests in
		// the order of the GPU kernel instantiation fmm_dualTraversal<true> (:504-542)
		stack.push_back(int2{0, 0});
		while (!stack.empty())
		{
			int2 np = stack.back();
			stack.pop_back();
			bool leaves = kd_lchild(np.x) >= ntot && kd_lchild(np.y) >= ntot;
			if (np.x == np.y && kd_lchild(np.x) < ntot)
			{
				stack.push_back({kd_lchild(np.x), kd_lchild(np.x)});
				stack.push_back({kd_lchild(np.x), kd_rchild(np.x)});
				stack.push_back({kd_rchild(np.x), kd_rchild(np.x)});
			}
			else if (np.x != np.y && kd_admissible(tree, np.x, np.y, radius))
				m2l_list.push_back(np);
			else if (leaves)
			{
				if (np.x != np.y)
					p2p_list.push_back(np);
			}
			else if (kd_lchild(np.x) >= ntot || (kd_lchild(np.y) < ntot
				&& kd_size(tree.lbound[np.x], tree.rbound[np.x]) <= kd_size(tree.lbound[np.y], tree.rbound[np.y])))
			{
				stack.push_back({np.x, kd_lchild(np.y)});
				stack.push_back({np.x, kd_rchild(np.y)});
			}
			else
			{
				stack.push_back({kd_lchild(np.x), np.y});
				stack.push_back({kd_rchild(np.x), np.y});
			}
		}
	}
	counts[0] = (long long)p2p_list.size();
	counts[1] = (long long)m2l_list.size();
	if ((long long)p2p_list.size() > cap || (long long)m2l_list.size() > cap)
		return -2;
	std::memcpy(p2p_pairs, p2p_list.data(), p2p_list.size()*sizeof(int2));
	std::memcpy(m2l_pairs, m2l_list.data(), m2l_list.size()*sizeof(int2));

	for (int i = 0; i < n; ++i)
		a[i] = VEC{};
	int list_n = p2p_list.size();
	if (coll)
	{
		int max_mlt = (n-1) / m + 1;
		fmm_p2p3_kdtree_cpu(a, tree, p, p2p_list.data(), &list_n, max_mlt, EPS2);
		fmm_p2p3_self_kdtree_cpu(a, tree, p, L, max_mlt, EPS2);
	}
	list_n = m2l_list.size();
	fmm_c2c3_kdtree_cpu(tree, m2l_list.data(), &list_n, EPS2);
	for (int l = 1; l <= L-1; ++l)
		fmm_pushl3_kdtree_cpu(tree, l);
	fmm_pushLeaves3_kdtree_cpu(a, p, tree, L);
	if (param != nullptr)
		rescale_cpu(a, n, param);
	return L;
}

// ---- reference GPU direct sum (generic SIMT source compiled for sm_100), for timing on a GPU box ----
double ref_direct3_gpu_seconds(const float *pos, float *acc, int n, const float *param6, int reps)
{
	VEC *d_p, *d_a; SCAL *d_par;
	if (cudaMalloc(&d_p, sizeof(VEC)*n) != cudaSuccess) return -1;
	cudaMalloc(&d_a, sizeof(VEC)*n);
	cudaMalloc(&d_par, sizeof(SCAL)*6);
	cudaMemcpy(d_p, pos, sizeof(VEC)*n, cudaMemcpyHostToDevice);
	cudaMemcpy(d_par, param6, sizeof(SCAL)*6, cudaMemcpyHostToDevice);
	direct3(d_p, d_a, n, d_par); // warm-up
	auto t0 = steady_clock::now();
	for (int r = 0; r < reps; ++r)
		direct3(d_p, d_a, n, d_par); // synchronises internally (direct.cuh:243-244)
	double s = duration_cast<microseconds>(steady_clock::now() - t0).count() * 1e-6 / reps;
	cudaMemcpy(acc, d_a, sizeof(VEC)*n, cudaMemcpyDeviceToHost);
	cudaFree(d_p); cudaFree(d_a); cudaFree(d_par);
	return s;
}

// ---- reference GPU kd-tree FMM under its own leapfrog (main3.cu:59-63,841-846; generic SIMT source compiled for
// sm_100): the second stated baseline of bench.py.  buf = [pos | vel | acc] on the host (9 n floats), read back at
// the end (tree order, like the reference's own snapshots).  Returns seconds per step or < 0 on a CUDA error. ----
double ref_fmm3_gpu_step_seconds(float *buf, int n, const float *param6, double dt, int warm, int steps)
{
	SCAL *d_buf, *d_par;
	if (cudaMalloc(&d_buf, sizeof(VEC)*3*(size_t)n) != cudaSuccess) return -1;
	if (cudaMalloc(&d_par, sizeof(SCAL)*6) != cudaSuccess) return -1;
	cudaMemcpy(d_buf, buf, sizeof(VEC)*2*(size_t)n, cudaMemcpyHostToDevice);
	cudaMemcpy(d_par, param6, sizeof(SCAL)*6, cudaMemcpyHostToDevice);
	SCAL dts = (SCAL)dt;
	compute_force(coulombOscillatorFMMKD3, d_buf, n, d_par); // main3.cu:835-839
	for (int s = 0; s < warm; ++s)
		leapfrog(coulombOscillatorFMMKD3, d_buf, n, d_par, dts, step, 1);
	if (cudaDeviceSynchronize() != cudaSuccess) return -2;
	auto t0 = steady_clock::now();
	for (int s = 0; s < steps; ++s)
		leapfrog(coulombOscillatorFMMKD3, d_buf, n, d_par, dts, step, 1);
	if (cudaDeviceSynchronize() != cudaSuccess) return -3;
	double sec = duration_cast<microseconds>(steady_clock::now() - t0).count() * 1e-6 / (steps > 0 ? steps : 1);
	cudaMemcpy(buf, d_buf, sizeof(VEC)*3*(size_t)n, cudaMemcpyDeviceToHost);
	cudaFree(d_buf); cudaFree(d_par);
	return sec;
}

// one evaluation of a reference GPU evaluator on host data (0 direct3, 1 fmm_cart3_kdtree with the current globals);
// pos (and vel behind it, when b_unsort is false) come back as the evaluator left them
int ref_eval_gpu(int which, float *buf, int n, const float *param6)
{
	SCAL *d_buf, *d_par;
	if (cudaMalloc(&d_buf, sizeof(VEC)*3*(size_t)n) != cudaSuccess) return -1;
	if (cudaMalloc(&d_par, sizeof(SCAL)*6) != cudaSuccess) return -1;
	cudaMemcpy(d_buf, buf, sizeof(VEC)*2*(size_t)n, cudaMemcpyHostToDevice);
	cudaMemcpy(d_par, param6, sizeof(SCAL)*6, cudaMemcpyHostToDevice);
	if (which == 0) compute_force(direct3, d_buf, n, d_par);
	else if (which == 1) compute_force(fmm_cart3_kdtree, d_buf, n, d_par);
	else return -4;
	if (cudaDeviceSynchronize() != cudaSuccess) return -2;
	cudaMemcpy(buf, d_buf, sizeof(VEC)*3*(size_t)n, cudaMemcpyDeviceToHost);
	cudaFree(d_buf); cudaFree(d_par);
	return 0;
}

} // extern "C"
