// host_io2.cpp -- host-side pieces of the 2D `nbco` surface (SCAL = double, VEC = double2) that never
// touch the GPU: initial conditions, default beam parameters, binary state file.  Byte-compatible
// with the reference (Simulation/main.cu): centerDist :97-106, adjustRMS :108-118, initKV :120-145,
// initGA :147-170, generator and discard :779-780, default beam :272,294-313, file layout = all
// positions then all velocities (main.cu usage text), n = bytes / 2 / sizeof(VEC).

#include "../../include/nbco.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

namespace nbco { void set_error(const char *fmt, ...); }

namespace {

struct V2 { double x, y; };
constexpr double kTwoPi = 6.283185307179586476925286766559; // constants.cuh twopi

void center_dist(V2 *d, int64_t n)
{
	V2 s{0.0, 0.0};
	for (int64_t i = 0; i < n; ++i) { s.x += d[i].x; s.y += d[i].y; }
	s.x /= (double)n; s.y /= (double)n;
	for (int64_t i = 0; i < n; ++i) { d[i].x -= s.x; d[i].y -= s.y; }
}

void adjust_rms(V2 *d, int64_t n, V2 adj)
{
	V2 s{0.0, 0.0};
	for (int64_t i = 0; i < n; ++i) { s.x += d[i].x*d[i].x; s.y += d[i].y*d[i].y; }
	s.x /= (double)n; s.y /= (double)n;
	s.x = std::sqrt(s.x); s.y = std::sqrt(s.y);
	const V2 f{adj.x / s.x, adj.y / s.y};
	for (int64_t i = 0; i < n; ++i) { d[i].x *= f.x; d[i].y *= f.y; }
}

std::mt19937_64 make_gen()
{
	std::mt19937_64 gen(5351550349027530206ULL);
	gen.discard(624*2);
	return gen;
}

} // namespace

extern "C" {

int nbco_init_ga2(double *h, int64_t n, const double *x2, const double *u2)
{
	if (!h || n <= 0 || !x2 || !u2) { nbco::set_error("bad argument"); return NBCO_ERR_INVALID; }
	std::mt19937_64 gen = make_gen();
	std::normal_distribution<double> dist(0.0, 1.0);
	V2 *data = reinterpret_cast<V2 *>(h);
	for (int64_t i = 0; i < 4*n; ++i) h[i] = dist(gen);
	for (int64_t i = 0; i < n; ++i) { data[i].x *= x2[0]; data[i].y *= x2[1]; }
	for (int64_t i = n; i < 2*n; ++i) { data[i].x *= u2[0]; data[i].y *= u2[1]; }
	center_dist(data, n); adjust_rms(data, n, V2{x2[0], x2[1]});
	center_dist(data + n, n); adjust_rms(data + n, n, V2{u2[0], u2[1]});
	return NBCO_OK;
}

int nbco_init_kv2(double *h, int64_t n, const double *A2, const double *om2)
{
	if (!h || n <= 0 || !A2 || !om2) { nbco::set_error("bad argument"); return NBCO_ERR_INVALID; }
	std::mt19937_64 gen = make_gen();
	std::uniform_real_distribution<double> dist(0.0, 1.0);
	V2 *data = reinterpret_cast<V2 *>(h);
	for (int64_t i = 0; i < n; ++i)
	{
		// three draws per particle in this order (main.cu:130)
		const double eta = dist(gen);
		const double etax = kTwoPi * dist(gen);
		const double etay = kTwoPi * dist(gen);
		const double rt = std::sqrt(eta), rt1 = std::sqrt(1 - eta);
		data[i].x = A2[0] * rt * std::cos(etax);
		data[i].y = A2[1] * rt1 * std::cos(etay);
		data[i + n].x = A2[0] * om2[0] * rt * std::sin(etax);
		data[i + n].y = A2[1] * om2[1] * rt1 * std::sin(etay);
	}
	center_dist(data, n); adjust_rms(data, n, V2{A2[0] / 2, A2[1] / 2});
	center_dist(data + n, n); adjust_rms(data + n, n, V2{om2[0] * A2[0] / 2, om2[1] * A2[1] / 2});
	return NBCO_OK;
}

int nbco_beam_params2(const double *omega0, const double *emit, double tune_dep_y, double *out5)
// r.m.s.-matched beam of main.cu:294-313 (quartic for the depressed phase advance in x)
{
	if (!omega0 || !emit || !out5) { nbco::set_error("bad argument"); return NBCO_ERR_INVALID; }
	V2 omega, domega, A;
	omega.y = tune_dep_y * omega0[1];
	A.y = 2 * std::sqrt(emit[1] / omega.y);
	const double A2 = A.y * A.y;
	domega.y = (omega0[1] + omega.y) * (omega0[1] - omega.y);
	const double om0x2 = omega0[0] * omega0[0], om0x4 = om0x2 * om0x2, om0x6 = om0x4 * om0x2;
	const double c = -2 * om0x2, d = -A2 * domega.y * domega.y / (4 * emit[0]), p = c, q = d;
	const double Delta0 = 16 * om0x4, Delta1 = 27 * d * d + 128 * om0x6;
	const double Q = std::cbrt((Delta1 + std::sqrt((27 * d * d + 256 * om0x6) * (27 * d * d))) / 2);
	const double S = std::sqrt((-2 * p + (Q + Delta0 / Q)) / 3) / 2;
	omega.x = S - std::sqrt(-4 * S * S - 2 * p - q / S) / 2; // sol[3]
	A.x = 2 * std::sqrt(emit[0] / omega.x);
	out5[0] = A.x; out5[1] = A.y; out5[2] = omega.x; out5[3] = omega.y;
	out5[4] = domega.y * A.y * (A.x + A.y) / 2; // xi
	return NBCO_OK;
}

int nbco_state_read2(const char *path, double **out, int64_t *n)
{
	FILE *f = fopen(path, "rb");
	if (!f) { nbco::set_error("Error: cannot read from input location."); return NBCO_ERR_INVALID; }
	fseek(f, 0, SEEK_END);
	long long bytes = ftell(f);
	fseek(f, 0, SEEK_SET);
	int64_t nb = bytes / 2 / (2 * (long long)sizeof(double));
	double *buf = (double *)malloc(sizeof(double) * 4 * (size_t)(nb > 0 ? nb : 1));
	size_t got = fread(buf, 1, sizeof(double) * 4 * (size_t)nb, f);
	fclose(f);
	if (got != sizeof(double) * 4 * (size_t)nb) { free(buf); nbco::set_error("short read"); return NBCO_ERR_INVALID; }
	*out = buf; *n = nb;
	return NBCO_OK;
}

int nbco_state_write2(const char *path, const double *h, int64_t n)
{
	FILE *f = fopen(path, "wb");
	if (!f) { nbco::set_error("Error: cannot write on output location."); return NBCO_ERR_INVALID; }
	size_t put = fwrite(h, 1, sizeof(double) * 4 * (size_t)n, f);
	fclose(f);
	return put == sizeof(double) * 4 * (size_t)n ? NBCO_OK : NBCO_ERR_INVALID;
}

} // extern "C"
