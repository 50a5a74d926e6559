// Minimal C++ host program over the C-ABI (no Python, no reference code): builds a synthetic
// self-gravitating disk, integrates it with RKF7(8) on the device and prints the throughput.
//
//   g++ -O2 -std=c++17 examples/disk_rkf78.cpp -Iinclude -Lsolaris_b200 -lsolaris_b200 \
//       -Wl,-rpath,'$ORIGIN/../solaris_b200' -o examples/disk_rkf78
//   ./examples/disk_rkf78 [bodies=65536] [steps=3] [Phases.dat to append snapshots to]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "solaris_b200.h"

int main(int argc, char **argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 65536;
	const int steps = argc > 2 ? atoi(argv[2]) : 3;
	const char *phases_path = argc > 3 ? argv[3] : nullptr;
	const double k2 = 2.959122082855911025e-4;
	std::mt19937_64 rng(20240601);
	std::uniform_real_distribution<double> U(0.0, 1.0);
	std::vector<double> y0(6 * (size_t)n, 0.0), mass(n, 0.0), zero(n, 0.0);
	std::vector<int> type(n, 4 /* ProtoPlanet */), mig(n, 0), id(n);
	mass[0] = 1.0; type[0] = 1;   // central body at the origin
	for (int i = 0; i < n; i++) id[i] = i + 1;
	for (int i = 1; i < n; i++) {
		// circular orbits, a ~ U(5,6) AU, small inclinations
		const double a = 5.0 + U(rng), phi = 2 * M_PI * U(rng), inc = 0.05 * (U(rng) - 0.5);
		mass[i] = (0.001 + 0.099 * U(rng)) / 3.3294605e5;
		const double v = std::sqrt(k2 * (1.0 + mass[i]) / a);
		double *p = &y0[6 * (size_t)i];
		p[0] = a * std::cos(phi); p[1] = a * std::sin(phi) * std::cos(inc); p[2] = a * std::sin(phi) * std::sin(inc);
		p[3] = -v * std::sin(phi); p[4] = v * std::cos(phi) * std::cos(inc); p[5] = v * std::cos(phi) * std::sin(inc);
	}
	sol_ctx *ctx = nullptr;
	if (sol_create(0, &ctx) != SOL_OK) { fprintf(stderr, "sol_create: %s\n", sol_last_error(nullptr)); return 1; }
	const int counts[7] = {1, 0, 0, n - 1, 0, 0, 0};
	if (sol_set_frame(ctx, 0) != SOL_OK ||
	    sol_set_bodies(ctx, counts, y0.data(), mass.data(), zero.data(), zero.data(), zero.data(), zero.data(), zero.data(),
	                   zero.data(), type.data(), mig.data(), id.data()) != SOL_OK ||
	    sol_set_nebula(ctx, nullptr) != SOL_OK) {
		fprintf(stderr, "setup: %s\n", sol_last_error(ctx));
		return 1;
	}
	double t = 0.0, h = 0.08, hdid = 0.0, info[4];
	double pairs = 0.0;
	const auto t0 = std::chrono::steady_clock::now();
	for (int s = 0; s < steps; s++) {
		if (sol_step(ctx, SOL_RUNGE_KUTTA_FEHLBERG78, &t, &h, &hdid, info) != SOL_OK) {
			fprintf(stderr, "sol_step: %s\n", sol_last_error(ctx));
			return 1;
		}
		pairs += info[3];
		// per-step event check on the device: only counts (and, if any, 120-byte records) come back
		int ev[3] = {0, 0, 0};
		if (sol_detect_events(ctx, /*ejection au*/ 50.0, /*hit centrum au*/ 0.1, /*collision factor*/ 0.0, ev) != SOL_OK) {
			fprintf(stderr, "sol_detect_events: %s\n", sol_last_error(ctx));
			return 1;
		}
		if (ev[0] + ev[1] > 0) printf("  %d ejections, %d hit centrums\n", ev[0], ev[1]);
		printf("step %d: t = %.6f d, hDid = %.6f d, attempts = %d, errorMax = %.3e\n", s, t, hdid, (int)info[0], info[1]);
	}
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	double integrals[16];
	sol_integrals(ctx, integrals);
	printf("%d bodies, %d RKF78 steps in %.3f s: %.3e pair interactions/s; total energy %.12e\n", n, steps, sec, pairs / sec, integrals[15]);
	// snapshot in the reference's Phases.dat format, assembled on the device
	if (phases_path != nullptr && sol_write_phases(ctx, phases_path, t) != SOL_OK) {
		fprintf(stderr, "sol_write_phases: %s\n", sol_last_error(ctx));
		return 1;
	}
	sol_destroy(ctx);
	return 0;
}
