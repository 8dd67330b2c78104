// host_io.cpp -- host-side pieces of the nbco3 surface that never touch the GPU:
// initial conditions and the binary state file.  Byte-compatible with the reference
// (Simulation/main3.cu): initGA :114-137 with centerDist :71-80 and adjustRMS :82-92,
// initU :94-112, generator and discard count :662-664, file layout :629-652,848-858.
// The sampler is libstdc++'s std::normal_distribution<float> over std::mt19937_64, as in
// the reference, so the stream of numbers is identical when built with the same libstdc++.

#include "../../include/nbco.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

namespace nbco { void set_error(const char *fmt, ...); }

namespace {

struct V3 { float x, y, z; };

void center_dist(V3 *d, int64_t n)
{
	V3 s{0.f, 0.f, 0.f};
	for (int64_t i = 0; i < n; ++i) { s.x += d[i].x; s.y += d[i].y; s.z += d[i].z; }
	s.x /= (float)n; s.y /= (float)n; s.z /= (float)n;
	for (int64_t i = 0; i < n; ++i) { d[i].x -= s.x; d[i].y -= s.y; d[i].z -= s.z; }
}

void adjust_rms(V3 *d, int64_t n, V3 adj)
{
	V3 s{0.f, 0.f, 0.f};
	for (int64_t i = 0; i < n; ++i) { s.x += d[i].x*d[i].x; s.y += d[i].y*d[i].y; s.z += d[i].z*d[i].z; }
	s.x /= (float)n; s.y /= (float)n; s.z /= (float)n;
	s.x = std::sqrt(s.x); s.y = std::sqrt(s.y); s.z = std::sqrt(s.z);
	V3 f{adj.x / s.x, adj.y / s.y, adj.z / s.z};
	for (int64_t i = 0; i < n; ++i) { d[i].x *= f.x; d[i].y *= f.y; d[i].z *= f.z; }
}

void init_ga(V3 *data, int64_t n, V3 x, V3 u, std::mt19937_64 &gen)
{
	std::normal_distribution<float> dist(0.f, 1.f);
	float *s = reinterpret_cast<float *>(data);
	for (int64_t i = 0; i < 2*n*3; ++i) s[i] = dist(gen);
	for (int64_t i = 0; i < n; ++i) { data[i].x *= x.x; data[i].y *= x.y; data[i].z *= x.z; }
	for (int64_t i = n; i < 2*n; ++i) { data[i].x *= u.x; data[i].y *= u.y; data[i].z *= u.z; }
	center_dist(data, n); adjust_rms(data, n, x);
	center_dist(data + n, n); adjust_rms(data + n, n, u);
}

} // namespace

extern "C" {

int nbco_init_ga(float *h, int64_t n, const float *sx, const float *su)
{
	if (!h || n <= 0 || !sx || !su) { nbco::set_error("bad argument"); return NBCO_ERR_INVALID; }
	std::mt19937_64 gen(5351550349027530206ULL);
	gen.discard(624*2);
	init_ga(reinterpret_cast<V3 *>(h), n, V3{sx[0], sx[1], sx[2]}, V3{su[0], su[1], su[2]}, gen);
	return NBCO_OK;
}

int nbco_init_test_cube(float *h, int64_t n, const float *sx, const float *su)
{
	if (!h || n <= 0 || !sx || !su) { nbco::set_error("bad argument"); return NBCO_ERR_INVALID; }
	std::mt19937_64 gen(5351550349027530206ULL);
	gen.discard(624*2);
	V3 *d = reinterpret_cast<V3 *>(h);
	init_ga(d, n, V3{sx[0], sx[1], sx[2]}, V3{su[0], su[1], su[2]}, gen);
	std::uniform_real_distribution<float> dx(-1.f, 1.f), dy(-1.f, 1.f), dz(-1.f, 1.f);
	for (int64_t i = 0; i < n; ++i) { d[i].x = dx(gen); d[i].y = dy(gen); d[i].z = dz(gen); }
	center_dist(d, n);
	return NBCO_OK;
}

int nbco_state_read(const char *path, float **out, int64_t *n)
{
	FILE *f = fopen(path, "rb");
	if (!f) { nbco::set_error("Error: cannot read from input location."); return NBCO_ERR_INVALID; }
	fseek(f, 0, SEEK_END);
	long long bytes = ftell(f);
	fseek(f, 0, SEEK_SET);
	int64_t nb = bytes / 2 / (3 * (long long)sizeof(float)); // n = bytes/2/sizeof(VEC), main3.cu:636
	float *buf = (float *)malloc(sizeof(float) * 6 * (size_t)(nb > 0 ? nb : 1));
	size_t got = fread(buf, 1, sizeof(float) * 6 * (size_t)nb, f);
	fclose(f);
	if (got != sizeof(float) * 6 * (size_t)nb) { free(buf); nbco::set_error("short read"); return NBCO_ERR_INVALID; }
	*out = buf; *n = nb;
	return NBCO_OK;
}

int nbco_state_write(const char *path, const float *h, int64_t n)
{
	FILE *f = fopen(path, "wb");
	if (!f) { nbco::set_error("Error: cannot write on output location."); return NBCO_ERR_INVALID; }
	size_t put = fwrite(h, 1, sizeof(float) * 6 * (size_t)n, f);
	fclose(f);
	return put == sizeof(float) * 6 * (size_t)n ? NBCO_OK : NBCO_ERR_INVALID;
}

void nbco_free(void *p) { free(p); }

} // extern "C"
