// nbco3 -- command-line driver with the surface of the reference's Simulation/main3.cu (C++20 host
// code above the C ABI of include/nbco.h; every number is computed by libnbco.so on the GPU).
//
// Kept from the reference: option names and defaults (main3.cu:229-246, 254-305, parsing :247-623),
// "-iters n" running n+1 iterations (:232,357), snapshots "<out>/out<iter>_<to_string(dt)>.bin" written
// when iter % steps == 0 after the step (:841-858), args.txt (:671-675), the parameter block
// {xi/N, 0, 0, w0x^2, w0y^2, w0z^2} (:685-692), the initial compute_force (:835-839), -test / -test2 /
// -accuracy (:737-831), error messages and the -1 exit code.  Differences: "-integ" accepts both "-fr" and
// "fr" (the reference skips the first character, :389-395); "-cpu", "-cpu-threads", "-cacheline" are
// rejected because this build has no CPU path; orders above NBCO_MAX_ORDER are rejected.

#include "../../include/nbco.h"
#include <cuda_runtime.h>
#include <chrono>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <thread>
#include <barrier>
#include <atomic>

namespace {

int fail(const std::string &msg)
{
	std::cerr << msg << std::endl;
	return -1;
}

#define CK(call) do { if ((call) != NBCO_OK) { std::cerr << nbco_last_error() << std::endl; return -1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::cerr << "GPUassert: " << cudaGetErrorString(e_) << ' ' << __FILE__ << ' ' << __LINE__ << std::endl; return -1; } } while (0)

const char *kHelp =
	"Usage: nbco3 [options] [input]\n"
	"Options:\n"
	"  -h, -help              print this help and exit\n"
	"  -o <dir>               output directory (must exist). Default is 'out'\n"
	"  -n <n>                 number of particles (ignored with an input file). Default is 30001\n"
	"  -ds <dt>               time step. Default is 5e-4\n"
	"  -iters <n>             number of iterations (n+1 steps are done). Default is 30000\n"
	"  -steps <n>             iterations between two snapshots. Default is 200\n"
	"  -integ <eu|fr|pefrl>   symplectic Euler, Forest-Ruth or PEFRL. Default is leapfrog\n"
	"  -p <order>             FMM order (1..10). Default is 3\n"
	"  -r <radius>            multipole acceptance parameter. Default is 1\n"
	"  -eps <eps>             softening length. Default is 1e-9\n"
	"  -i <x>                 density inhomogeneity factor for the tree depth. Default is 1\n"
	"  -maxlevel <L>          force the tree depth. Default is automatic\n"
	"  -ncoll                 skip the near field (P2P)\n"
	"  -tree-steps <k>        (extension) rebuild the kd-tree every k force evaluations. Default is 8\n"
	"  -m2l-first <0|1>       (extension) 1: MAC before leaf test (reference GPU order, default), 0: reference CPU order\n"
	"  -gpus <g>              (extension) partition the kd-tree over g GPUs of this node (1, 2, 4 or 8; NVLink peer memory)\n"
	"  -reproducible          (extension) interaction lists bucketed by target: bit-reproducible forces, no float atomics\n"
	"  -accuracy <v>          search (p, r) for a mean relative error below v, then run\n"
	"  -test                  timing and error against the direct sum for p = 1..10 on a uniform cube\n"
	"  -test2                 error against the direct sum over tree_steps+1 Euler steps\n"
	"  -xi <v>                perveance-like coupling. Default is 2e-6\n"
	"  -omega0 <x> <y>        transverse oscillator frequencies (z stays 1)\n"
	"  -x <sx> <sy> <sz>      position standard deviations. Default is 0.003 0.001 0.01\n"
	"  -u <ux> <uy> <uz>      velocity standard deviations. Default is omega0 * x\n";

struct Device
{
	nbco_ctx *ctx = nullptr;
	float *buf = nullptr, *par = nullptr, *tmp = nullptr;
	~Device() { if (buf) cudaFree(buf); if (par) cudaFree(par); if (tmp) cudaFree(tmp); if (ctx) nbco_destroy(ctx); }
};

// The main loop of main3.cu:835-858 over g GPUs of one node: one host thread and one context per GPU, rank r owns the
// subtree of kd node (log2 g, r) (csrc/peer.cu).  The contexts live in one process, so they are
// wired with nbco_peer_attach_local (no CUDA IPC); every rank holds its own copy of the state buffer, steps its own
// range and the ranks meet inside the evaluator.  Snapshots: nbco_peer_gather, rank 0 writes the file.
int run_multi_gpu(int gpus, nbco_config cfg, int scheme, std::vector<float> &host, int64_t n, const float *par6, float dt,
                  long nIters, long nSteps, const std::string &strout)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < gpus) return fail("Error: -gpus " + std::to_string(gpus) + " but " + std::to_string(ndev) + " CUDA devices are visible.");
	if (gpus > 8 || (gpus & (gpus - 1))) return fail("Error: -gpus must be 1, 2, 4 or 8.");
	const size_t vb = sizeof(float) * 3 * (size_t)n;
	std::vector<Device> dev(gpus);
	std::barrier sync(gpus);
	std::atomic<int> failed{0};
	std::vector<std::string> errs(gpus);
	auto rank_main = [&](int r)
	{
		auto bail = [&](const std::string &m) { errs[r] = m; failed.store(1); };
#define RK(call) do { if (!failed.load() && (call) != NBCO_OK) bail(nbco_last_error()); } while (0)
#define RU(call) do { if (!failed.load()) { cudaError_t e_ = (call); if (e_ != cudaSuccess) bail(std::string("GPUassert: ") + cudaGetErrorString(e_)); } } while (0)
		Device &d = dev[r];
		nbco_config c = cfg;
		c.device = r; c.rank = r; c.world = gpus; c.unsort = 0;
		RU(cudaSetDevice(r));
		RK(nbco_create(&c, &d.ctx));
		RU(cudaMalloc(&d.buf, 3 * vb));
		RU(cudaMalloc(&d.par, 6 * sizeof(float)));
		RU(cudaMemcpy(d.buf, host.data(), 2 * vb, cudaMemcpyHostToDevice));
		RU(cudaMemcpy(d.par, par6, 6 * sizeof(float), cudaMemcpyHostToDevice));
		unsigned char handles[192];
		RK(nbco_peer_export(d.ctx, n, handles));
		sync.arrive_and_wait();
		for (int q = 0; q < gpus && !failed.load(); ++q)
			if (q != r) RK(nbco_peer_attach_local(d.ctx, q, dev[q].ctx));
		RK(nbco_peer_commit(d.ctx));
		sync.arrive_and_wait();
		if (failed.load()) return; // every rank sees the flag after the barrier: nobody enters a collective call alone
		RK(nbco_compute_force(d.ctx, NBCO_EVAL_COULOMB_FMM3_KD, d.buf, n, d.par));
		long iter = 0;
		while (iter < nIters)
		{
			sync.arrive_and_wait();
			if (failed.load()) return;
			const long next = (iter % nSteps == 0) ? iter : (iter / nSteps + 1) * nSteps;
			const long todo = std::min(next, nIters - 1) - iter + 1;
			RK(nbco_integrate(d.ctx, scheme, NBCO_EVAL_COULOMB_FMM3_KD, d.buf, n, d.par, dt, todo));
			iter += todo;
			if ((iter - 1) % nSteps == 0)
			{
				sync.arrive_and_wait();
				if (failed.load()) return;
				RK(nbco_peer_gather(d.ctx, d.buf, n));
				if (r == 0 && !failed.load())
				{
					std::cout << (iter - 1) << ' ' << std::flush;
					RU(cudaMemcpy(host.data(), d.buf, 2 * vb, cudaMemcpyDeviceToHost));
					const std::string name = strout + "/out" + std::to_string(iter - 1) + '_' + std::to_string(dt) + ".bin";
					if (!failed.load() && nbco_state_write(name.c_str(), host.data(), n) != NBCO_OK)
						bail("Error: cannot write on output location. Check that \"" + strout + "\" folder exists. Create it if not.");
				}
			}
		}
		sync.arrive_and_wait();
		if (d.ctx) nbco_peer_detach(d.ctx);
		sync.arrive_and_wait(); // every rank has closed its view of the others before anybody frees (Device destructors)
#undef RK
#undef RU
	};
	std::vector<std::thread> th;
	for (int r = 0; r < gpus; ++r) th.emplace_back(rank_main, r);
	for (auto &t : th) t.join();
	std::cout << std::endl;
	if (failed.load())
	{
		for (const auto &e : errs) if (!e.empty()) std::cerr << e << std::endl;
		return -1;
	}
	return 0;
}

} // namespace

int main(int argc, const char **argv)
{
	std::cout << "N-body coulomb oscillators (B200-native hot path; surface of nbco3, Copyright (C) 2021-24 Alessandro Lo Cuoco)\n\n"
	             "Type 'nbco3 -h' for a brief documentation.\n\n";
	int64_t nBodies = 30001;
	float dt = 5.e-4f;
	long nIters = 30001, nSteps = 200;
	std::string strout("out"), strin;
	bool in = false, test = false, test2 = false, b_accuracy = false;
	int gpus = 1;
	float accuracy = 0.001f, xi = 2.e-6f;
	float omega0[3] = {1.095f, 1.0f, 1.0f}, x[3] = {0.003f, 0.001f, 0.01f}, u[3];
	for (int k = 0; k < 3; ++k) u[k] = omega0[k] * x[k]; // computed before parsing like main3.cu:241-245
	int scheme = NBCO_LEAPFROG;
	nbco_config cfg;
	nbco_default_config(&cfg);

	auto need = [&](int i, int k) { return i + k < argc; };
	for (int i = 1; i < argc; ++i)
	{
		std::string a = argv[i];
		if (a.empty() || a[0] != '-') { strin = a; in = true; continue; }
		if (a == "-h" || a == "-help") { std::cout << kHelp; return 0; }
		else if (a == "-o") { if (!need(i, 1)) return fail("Error: no output location specified."); strout = argv[++i]; }
		else if (a == "-n") { if (!need(i, 1)) return fail("Error: no number of particles specified."); nBodies = atoll(argv[++i]); }
		else if (a == "-ds") { if (!need(i, 1)) return fail("Error: no time step specified."); dt = (float)atof(argv[++i]); }
		else if (a == "-iters") { if (!need(i, 1)) return fail("Error: no number of iterations specified."); nIters = atol(argv[++i]) + 1; }
		else if (a == "-steps") { if (!need(i, 1)) return fail("Error: no number of steps specified."); nSteps = atol(argv[++i]); }
		else if (a == "-integ")
		{
			if (!need(i, 1)) return fail("Error: no integrator specified.");
			std::string v = argv[++i];
			if (!v.empty() && v[0] == '-') v = v.substr(1);
			if (v == "eu") scheme = NBCO_EULER;
			else if (v == "fr") scheme = NBCO_FORESTRUTH;
			else if (v == "pefrl") scheme = NBCO_PEFRL;
			else return fail("Error: integrator not recognized.");
		}
		else if (a == "-p") { if (!need(i, 1)) return fail("Error: no FMM order specified."); cfg.order = atoi(argv[++i]); }
		else if (a == "-r") { if (!need(i, 1)) return fail("Error: no radius specified."); cfg.radius = (float)atof(argv[++i]); }
		else if (a == "-eps") { if (!need(i, 1)) return fail("Error: no softening specified."); float e = (float)atof(argv[++i]); cfg.eps2 = e * e; }
		else if (a == "-i") { if (!need(i, 1)) return fail("Error: no inhomogeneity specified."); cfg.dens_inhom = (float)atof(argv[++i]); }
		else if (a == "-maxlevel") { if (!need(i, 1)) return fail("Error: no level specified."); cfg.max_level = atoi(argv[++i]); }
		else if (a == "-ncoll") cfg.coll = 0;
		// extensions (not in the reference CLI): its global tree_steps (constants.cuh:45) and the traversal order
		else if (a == "-tree-steps") { if (!need(i, 1)) return fail("Error: no tree_steps specified."); cfg.tree_steps = atoi(argv[++i]); }
		else if (a == "-m2l-first") { if (!need(i, 1)) return fail("Error: no value specified."); cfg.m2l_first = atoi(argv[++i]); }
		else if (a == "-gpus") { if (!need(i, 1)) return fail("Error: no number of GPUs specified."); gpus = atoi(argv[++i]); if (gpus < 1) return fail("Error: invalid argument to '-gpus'"); }
		else if (a == "-reproducible") cfg.reproducible = 1;
		else if (a == "-accuracy") { if (!need(i, 1)) return fail("Error: no accuracy specified."); accuracy = (float)atof(argv[++i]); b_accuracy = true; }
		else if (a == "-test") test = true;
		else if (a == "-test2") test2 = true;
		else if (a == "-xi") { if (!need(i, 1)) return fail("Error: no xi specified."); xi = (float)atof(argv[++i]); }
		else if (a == "-omega0") { if (!need(i, 2)) return fail("Error: omega0 needs two values."); omega0[0] = (float)atof(argv[++i]); omega0[1] = (float)atof(argv[++i]); }
		else if (a == "-x") { if (!need(i, 3)) return fail("Error: -x needs three values."); for (int k = 0; k < 3; ++k) x[k] = (float)atof(argv[++i]); }
		else if (a == "-u") { if (!need(i, 3)) return fail("Error: -u needs three values."); for (int k = 0; k < 3; ++k) u[k] = (float)atof(argv[++i]); }
		else if (a == "-cpu" || a == "-cpu-threads" || a == "-cacheline")
			return fail("Error: this build has no CPU path (" + a + "); the reference's CPU path lives in the reference.");
		else return fail("Error: unrecognised option " + a);
	}
	if (nBodies < 8) return fail("Error: at least 8 particles are needed.");
	if (nSteps < 1) nSteps = 1;

	std::vector<float> host;
	if (in)
	{
		float *p = nullptr;
		int64_t n = 0;
		if (nbco_state_read(strin.c_str(), &p, &n) != NBCO_OK) return fail(nbco_last_error());
		nBodies = n;
		host.assign(p, p + 6 * n);
		nbco_free(p);
	}
	else
	{
		host.resize(6 * (size_t)nBodies);
		if (test) CK(nbco_init_test_cube(host.data(), nBodies, x, u));
		else CK(nbco_init_ga(host.data(), nBodies, x, u));
	}
	if (!test && !test2)
	{
		std::ofstream farg(strout + "/args.txt", std::ios::out);
		if (!farg)
			return fail("Error: cannot write on output location. Check that \"" + strout + "\" folder exists. Create it if not.");
		for (int i = 0; i < argc; ++i) farg << argv[i] << ' ';
	}
	const float par[6] = {xi / (float)nBodies, 0, 0, omega0[0] * omega0[0], omega0[1] * omega0[1], omega0[2] * omega0[2]};

	if (gpus > 1)
	{
		if (test || test2 || b_accuracy) return fail("Error: -test, -test2 and -accuracy run on one GPU (drop -gpus).");
		return run_multi_gpu(gpus, cfg, scheme, host, nBodies, par, dt, nIters, nSteps, strout);
	}

	Device d;
	const int64_t n = nBodies;
	const size_t vb = sizeof(float) * 3 * (size_t)n;
	CK(nbco_create(&cfg, &d.ctx));
	CU(cudaMalloc(&d.buf, 3 * vb));
	CU(cudaMalloc(&d.tmp, vb));
	CU(cudaMalloc(&d.par, sizeof(par)));
	CU(cudaMemcpy(d.buf, host.data(), 2 * vb, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d.par, par, sizeof(par), cudaMemcpyHostToDevice));
	float *d_acc = d.buf + 6 * n;

	auto set = [&](int order, float radius, int unsort) -> int { cfg.order = order; cfg.radius = radius; cfg.unsort = unsort; return nbco_set_config(d.ctx, &cfg); };
	// test_accuracy (main3.cu:139-182): mean rel_diff1 of the FMM against direct3, input order kept
	auto test_accuracy = [&](double &err) -> int
	{
		if (nbco_force_direct3(d.ctx, d.buf, d.tmp, n, d.par)) return -1;
		if (nbco_force_fmm3_kd(d.ctx, d.buf, d_acc, n, d.par)) return -1;
		return nbco_mean_rel_err(d.ctx, d_acc, d.tmp, n, &err, nullptr);
	};
	// test_time (main3.cu:707-735): warm-up, then doubling loop until min_loop seconds
	auto test_time = [&](double min_loop, double &sec) -> int
	{
		if (nbco_force_fmm3_kd(d.ctx, d.buf, d_acc, n, d.par)) return -1;
		int loop_n = 1, count = 0;
		auto t0 = std::chrono::steady_clock::now();
		double dur;
		do
		{
			for (int i = 0; i < loop_n; ++i) if (nbco_force_fmm3_kd(d.ctx, d.buf, d_acc, n, d.par)) return -1;
			count += loop_n; loop_n *= 2;
			dur = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		} while (dur < min_loop);
		sec = dur / count;
		return 0;
	};

	if (b_accuracy)
	{
		const int search_p[] = {1, 2, 3, 4, 5, 6};
		const float search_r[] = {1.11f, 1.25f, 1.43f, 1.67f, 2.f, 2.5f, 3.f};
		double best_time = FLT_MAX, best_acc = 0;
		float best_r = 1;
		int best_p = 3;
		cfg.coll = 1;
		std::cout << "Parameter optimization in progress, please wait" << std::flush;
		for (float r : search_r)
			for (int p : search_p)
			{
				double err, sec;
				if (set(p, r, 1) || test_accuracy(err)) return fail(nbco_last_error());
				if (err < accuracy)
				{
					if (test_time(0, sec)) return fail(nbco_last_error());
					if (sec < best_time) { best_r = r; best_p = p; best_acc = err; best_time = sec; }
				}
				std::cout << '.' << std::flush;
			}
		if (best_time == FLT_MAX) { std::cout << "\nOptimization failed!" << std::endl; return -1; }
		cfg.order = best_p; cfg.radius = best_r;
		std::cout << "\nBest parameters: r = " << best_r << ", p = " << best_p << ", time = " << best_time
		          << ", error = " << best_acc << std::endl;
	}

	if (test)
	{
		double sec, err;
		if (set(cfg.order, cfg.radius, 0) || test_time(1, sec)) return fail(nbco_last_error());
		std::cout << cfg.order << ": Average time: " << sec << " [s]" << std::endl;
		// test_time left the positions in tree order; that is a permutation of the same system
		for (int p = 1; p <= NBCO_MAX_ORDER; ++p)
		{
			if (set(p, cfg.radius, 1) || test_accuracy(err)) return fail(nbco_last_error());
			std::cout << p << ": Relative error: " << err << std::endl;
		}
	}
	else if (test2)
	{
		if (set(cfg.order, cfg.radius, 0)) return fail(nbco_last_error());
		for (int i = 0; i < cfg.tree_steps + 1; ++i)
		{
			double err;
			// FMM first (it may reorder pos/vel), then the direct sum on the same order
			if (nbco_force_fmm3_kd(d.ctx, d.buf, d.tmp, n, d.par)) return fail(nbco_last_error());
			if (nbco_force_direct3(d.ctx, d.buf, d_acc, n, d.par)) return fail(nbco_last_error());
			if (nbco_mean_rel_err(d.ctx, d.tmp, d_acc, n, &err, nullptr)) return fail(nbco_last_error());
			// pre_symplectic_euler(add_elastic, ...) (main3.cu:820-826, integrator.cuh:50-66): add_elastic SUBTRACTS k x from the
			// acc buffer, which test_accuracy left holding the direct-sum Coulomb field (main3.cu:172-173): the particles step
			// under Coulomb + elastic force:  a -= k x; v += a dt; x += v dt
			CK(nbco_add_elastic(d.ctx, d.buf, d_acc, n, d.par + 3));
			CK(nbco_step(d.ctx, d.buf + 3 * n, d_acc, dt, n));
			CK(nbco_step(d.ctx, d.buf, d.buf + 3 * n, dt, n));
			std::cout << "Relative error after " << i << " steps: " << err << std::endl;
		}
	}
	else
	{
		if (set(cfg.order, cfg.radius, 0)) return fail(nbco_last_error());
		CK(nbco_compute_force(d.ctx, NBCO_EVAL_COULOMB_FMM3_KD, d.buf, n, d.par));
		long iter = 0;
		while (iter < nIters)
		{
			// iterations up to and including the next snapshot iteration (iter % nSteps == 0)
			long next = (iter % nSteps == 0) ? iter : (iter / nSteps + 1) * nSteps;
			long todo = std::min(next, nIters - 1) - iter + 1;
			CK(nbco_integrate(d.ctx, scheme, NBCO_EVAL_COULOMB_FMM3_KD, d.buf, n, d.par, dt, todo));
			iter += todo;
			if ((iter - 1) % nSteps == 0)
			{
				std::cout << (iter - 1) << ' ' << std::flush;
				CU(cudaMemcpy(host.data(), d.buf, 2 * vb, cudaMemcpyDeviceToHost));
				std::string name = strout + "/out" + std::to_string(iter - 1) + '_' + std::to_string(dt) + ".bin";
				if (nbco_state_write(name.c_str(), host.data(), n) != NBCO_OK)
					return fail("Error: cannot write on output location. Check that \"" + strout + "\" folder exists. Create it if not.");
			}
		}
		std::cout << std::endl;
	}
	return 0;
}
