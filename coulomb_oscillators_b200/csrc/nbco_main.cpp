// nbco -- command-line driver with the surface of the reference's 2D program Simulation/main.cu (SCAL = double,
// DIM = 2; C++20 host code above the C ABI of include/nbco.h, every number is computed by libnbco.so on the GPU).
//
// Kept from the reference: option names and defaults (main.cu:262-313: n = 30001, ds = 5e-4, iters = 30000 (+1),
// steps = 200, p = 5 (:45), KV beam unless -ga, omega0 = (6.22, 6.21) * 2 pi, emittances (0.03e-3, 0.01e-3), the
// r.m.s.-matched default beam :294-313), "-x" sets A = 2x and "-A" sets x = A/2 (:652-703), "-u" makes omega = u/x and
// "-omega" makes u = omega x after parsing (:683,734-737), the parameter block {xi/N, 0, w0x^2, w0y^2} (:803-808), the
// initial compute_force (:863-867), snapshots "<out>/out<iter>_<to_string(dt)>.bin" when iter % steps == 0 (:876-879),
// args.txt (:788-792), -test (one timed evaluation, then the mean relative error against the direct sum for
// p = 1..10, :816-851), error texts and the -1 exit code.  Differences: "-cpu", "-cpu-threads", "-cacheline" are
// rejected (no CPU path); "-gpu" and "-gridsize" are accepted and ignored (launch shapes are chosen per kernel);
// "-integ" accepts both "-fr" and "fr"; main.cu itself does not compile at HEAD.

#include "../../include/nbco.h"
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

namespace {

int fail(const std::string &msg)
{
	std::cerr << msg << std::endl;
	return -1;
}

#define CK(call) do { if ((call) != NBCO_OK) { std::cerr << nbco_last_error() << std::endl; return -1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::cerr << "GPUassert: " << cudaGetErrorString(e_) << ' ' << __FILE__ << ' ' << __LINE__ << std::endl; return -1; } } while (0)

const char *kHelp =
	"Usage: nbco [options] [input]\n\n"
	"  [input] is the path to a file that contains a state of the system: the positions of all particles and\n"
	"  then their velocities in the same order (raw fp64). Without it the system is sampled from a KV or a\n"
	"  gaussian distribution.\n\n"
	"Other options:\n"
	"  -h or -help       Display this documentation.\n"
	"  -o <output>       Output folder (must exist). Default is './out'.\n"
	"  -n <npart>        Number of particles. Default is 30001. Ignored with [input].\n"
	"  -ds <v>           Time step. Default is 5e-4.\n"
	"  -iters <n>        Number of total simulation iterations. Default is 30000.\n"
	"  -steps <n>        Number of steps before saving to file. Default is 200.\n"
	"  -integ <name>     eu, fr or pefrl instead of the leapfrog.\n"
	"  -p <order>        FMM expansion order (1..10). Default is 5.\n"
	"  -r <radius>       Interaction radius. Must be 1 or greater. Default is 1.\n"
	"  -eps <v>          Smoothing factor. Must be greater than 0. Default is 1e-9.\n"
	"  -i <v>            A factor so that max FMM level is round(log(n*i/p^(3/2))). Default is 1.\n"
	"  -ncoll            P2P pass will not be calculated.\n"
	"  -gpu <blocksize>  Accepted and ignored (launch shapes are chosen per kernel).\n"
	"  -gridsize <n>     Accepted and ignored.\n"
	"  -test             Show relative errors and execution times of a single iteration.\n"
	"  -ga               Gaussian instead of KV distribution.\n"
	"  -xi <v>           Set the perveance.\n"
	"  -omega0 <vx> <vy> Set the phase advances.\n"
	"  -x <vx> <vy>      Set the std.dev. of positions.\n"
	"  -u <vx> <vy>      Set the std.dev. of velocities.\n"
	"  -A <vx> <vy>      Set the system semi-axes.\n"
	"  -omega <vx> <vy>  Set the depressed phase advances.\n"
	"Note: x = A / 2 and u = omega * A / 2.\n";

struct Device
{
	nbco_ctx *ctx = nullptr;
	double *buf = nullptr, *par = nullptr, *tmp = nullptr;
	~Device() { if (buf) cudaFree(buf); if (par) cudaFree(par); if (tmp) cudaFree(tmp); if (ctx) nbco_destroy(ctx); }
};

} // namespace

int main(int argc, const char **argv)
{
	std::cout << "N-body coulomb oscillators, 2D (B200-native hot path; surface of nbco, Copyright (C) 2021-24 Alessandro Lo Cuoco)\n\n"
	             "Type 'nbco -h' for a brief documentation.\n\n";
	int64_t nBodies = 30001;
	double dt = 5.e-4;
	long nIters = 30001, nSteps = 200;
	std::string strout("out"), strin;
	bool in = false, test = false, ga = false, calc_u = false, calc_omega = false;
	const double twopi = 6.283185307179586476925286766559;
	double omega0[2] = {6.22 * twopi, 6.21 * twopi}, emit[2] = {0.03e-3, 0.01e-3};
	double beam[5];
	if (nbco_beam_params2(omega0, emit, 0.8, beam) != NBCO_OK) return fail(nbco_last_error());
	double A[2] = {beam[0], beam[1]}, omega[2] = {beam[2], beam[3]}, xi = beam[4];
	double x[2] = {A[0] / 2, A[1] / 2}, u[2] = {omega[0] * A[0] / 2, omega[1] * A[1] / 2};
	int scheme = NBCO_LEAPFROG;
	nbco_config cfg;
	nbco_default_config(&cfg);
	cfg.order = 5; // main.cu:45

	auto need = [&](int i, int k) { return i + k < argc; };
	auto two = [&](int &i, double *v, const char *name) -> bool
	{
		if (!need(i, 2)) { std::cerr << "Error: missing argument(s) to '" << name << "'\n"; return false; }
		v[0] = atof(argv[i + 1]); v[1] = atof(argv[i + 2]);
		if (v[0] < 0 || v[1] < 0) { std::cerr << "Error: invalid argument(s) to '" << name << "': " << argv[i + 1] << ' ' << argv[i + 2] << '\n'; return false; }
		i += 2;
		return true;
	};
	for (int i = 1; i < argc; ++i)
	{
		std::string a = argv[i];
		if (a.empty() || a[0] != '-') { strin = a; in = true; continue; }
		if (a == "-h" || a == "-help") { std::cout << kHelp; return 0; }
		else if (a == "-o") { if (!need(i, 1)) return fail("Error: missing argument to '-o'"); strout = argv[++i]; }
		else if (a == "-n") { if (!need(i, 1)) return fail("Error: missing argument to '-n'"); nBodies = atoll(argv[++i]); if (nBodies < 1) return fail("Error: invalid argument to '-n'"); }
		else if (a == "-ds") { if (!need(i, 1)) return fail("Error: missing argument to '-ds'"); dt = atof(argv[++i]); }
		else if (a == "-iters") { if (!need(i, 1)) return fail("Error: missing argument to '-iters'"); nIters = atol(argv[++i]) + 1; }
		else if (a == "-steps") { if (!need(i, 1)) return fail("Error: missing argument to '-steps'"); nSteps = atol(argv[++i]); }
		else if (a == "-integ")
		{
			if (!need(i, 1)) return fail("Error: missing argument to '-integ'");
			std::string v = argv[++i];
			if (!v.empty() && v[0] == '-') v = v.substr(1);
			if (v == "eu") scheme = NBCO_EULER;
			else if (v == "fr") scheme = NBCO_FORESTRUTH;
			else if (v == "pefrl") scheme = NBCO_PEFRL;
			else return fail("Error: invalid argument to '-integ': " + v);
		}
		else if (a == "-p") { if (!need(i, 1)) return fail("Error: missing argument to '-p'"); cfg.order = atoi(argv[++i]); if (cfg.order < 1 || cfg.order > NBCO2_MAX_ORDER) return fail("Error: invalid argument to '-p'"); }
		else if (a == "-r") { if (!need(i, 1)) return fail("Error: missing argument to '-r'"); cfg.radius = (float)atof(argv[++i]); if (cfg.radius < 1) return fail("Error: invalid argument to '-r'"); }
		else if (a == "-eps") { if (!need(i, 1)) return fail("Error: missing argument to '-eps'"); double e = atof(argv[++i]); if (!(e > 0)) return fail("Error: invalid argument to '-eps'"); cfg.eps2_d = e * e; cfg.eps2 = (float)(e * e); }
		else if (a == "-i") { if (!need(i, 1)) return fail("Error: missing argument to '-i'"); cfg.dens_inhom = (float)atof(argv[++i]); if (!(cfg.dens_inhom > 0)) return fail("Error: invalid argument to '-i'"); }
		else if (a == "-ncoll") cfg.coll = 0;
		else if (a == "-gpu" || a == "-gridsize") { if (!need(i, 1)) return fail("Error: missing argument to '" + a + "'"); ++i; }
		else if (a == "-test") test = true;
		else if (a == "-ga") ga = true;
		else if (a == "-xi") { if (!need(i, 1)) return fail("Error: missing argument to '-xi'"); xi = atof(argv[++i]); if (xi < 0) return fail("Error: invalid argument to '-xi'"); }
		else if (a == "-omega0") { if (!two(i, omega0, "-omega0")) return -1; }
		else if (a == "-x") { if (!two(i, x, "-x")) return -1; A[0] = 2 * x[0]; A[1] = 2 * x[1]; }
		else if (a == "-u") { if (!two(i, u, "-u")) return -1; calc_omega = true; }
		else if (a == "-A") { if (!two(i, A, "-A")) return -1; x[0] = A[0] / 2; x[1] = A[1] / 2; }
		else if (a == "-omega") { if (!two(i, omega, "-omega")) return -1; calc_u = true; }
		else if (a == "-cpu" || a == "-cpu-threads" || a == "-cacheline")
			return fail("Error: this build has no CPU path (" + a + "); the reference's CPU path lives in the reference.");
		else return fail("Error: unrecognised option: " + a);
	}
	if (calc_omega) { omega[0] = u[0] / x[0]; omega[1] = u[1] / x[1]; }
	else if (calc_u) { u[0] = omega[0] * x[0]; u[1] = omega[1] * x[1]; }
	if (nSteps < 1) nSteps = 1;

	std::vector<double> host;
	if (in)
	{
		double *p = nullptr;
		int64_t n = 0;
		if (nbco_state_read2(strin.c_str(), &p, &n) != NBCO_OK) return fail(nbco_last_error());
		nBodies = n;
		host.assign(p, p + 4 * n);
		nbco_free(p);
	}
	else
	{
		std::cout << "perveance: " << xi << std::endl;
		std::cout << "dep. phase adv.: " << omega[0] << ' ' << omega[1] << std::endl;
		host.resize(4 * (size_t)nBodies);
		if (ga) CK(nbco_init_ga2(host.data(), nBodies, x, u));
		else CK(nbco_init_kv2(host.data(), nBodies, A, omega));
	}
	if (!test)
	{
		std::ofstream farg(strout + "/args.txt", std::ios::out);
		if (!farg)
			return fail("Error: cannot write on output location. Check that \"" + strout + "\" folder exists. Create it if not.");
		for (int i = 0; i < argc; ++i) farg << argv[i] << ' ';
	}
	const double par[4] = {xi / (double)nBodies, 0, omega0[0] * omega0[0], omega0[1] * omega0[1]};

	Device d;
	const int64_t n = nBodies;
	const size_t vb = sizeof(double) * 2 * (size_t)n;
	CK(nbco_create(&cfg, &d.ctx));
	CU(cudaMalloc(&d.buf, 3 * vb));
	CU(cudaMalloc(&d.tmp, vb));
	CU(cudaMalloc(&d.par, sizeof(par)));
	CU(cudaMemcpy(d.buf, host.data(), 2 * vb, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d.par, par, sizeof(par), cudaMemcpyHostToDevice));
	double *d_acc = d.buf + 4 * n;

	if (test)
	{
		CK(nbco_compute_force2(d.ctx, NBCO_EVAL_FMM2, d.buf, n, d.par)); // warming up (main.cu:818-822)
		auto t0 = std::chrono::steady_clock::now();
		CK(nbco_compute_force2(d.ctx, NBCO_EVAL_FMM2, d.buf, n, d.par));
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		std::cout << "Time elapsed: " << sec << " [s]" << std::endl;
		for (int p = 1; p <= NBCO2_MAX_ORDER; ++p)
		{
			cfg.order = p;
			CK(nbco_set_config(d.ctx, &cfg));
			double err;
			// test_accuracy (main.cu:172-216): FMM first (it sorts pos / vel), the direct sum on the same order
			CK(nbco_force_fmm2(d.ctx, d.buf, d.tmp, n, d.par));
			CK(nbco_force_direct2(d.ctx, d.buf, d_acc, n, d.par));
			CK(nbco_mean_rel_err2(d.ctx, d.tmp, d_acc, n, &err, nullptr));
			std::cout << p << ": Relative error: " << err << std::endl;
		}
		return 0;
	}
	CK(nbco_compute_force2(d.ctx, NBCO_EVAL_COULOMB_FMM2, d.buf, n, d.par));
	long iter = 0;
	while (iter < nIters)
	{
		long next = (iter % nSteps == 0) ? iter : (iter / nSteps + 1) * nSteps;
		long todo = std::min(next, nIters - 1) - iter + 1;
		CK(nbco_integrate2(d.ctx, scheme, NBCO_EVAL_COULOMB_FMM2, d.buf, n, d.par, dt, todo));
		iter += todo;
		if ((iter - 1) % nSteps == 0)
		{
			std::cout << (iter - 1) << ' ' << std::flush;
			CU(cudaMemcpy(host.data(), d.buf, 2 * vb, cudaMemcpyDeviceToHost));
			std::string name = strout + "/out" + std::to_string(iter - 1) + '_' + std::to_string(dt) + ".bin";
			if (nbco_state_write2(name.c_str(), host.data(), n) != NBCO_OK)
				return fail("Error: cannot write on output location. Check that \"" + strout + "\" folder exists. Create it if not.");
		}
	}
	std::cout << std::endl;
	return 0;
}
