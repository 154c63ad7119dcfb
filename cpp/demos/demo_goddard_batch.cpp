// cpp/demos/demo_goddard_batch.cpp -- what the engine is for: many perturbed instances of the
// reference's Goddard problem (tests/testGoddard.cpp:28-99: free final time, M = 6 shooting segments,
// xtol 1e-6, KD = 0, mu2 = 1) solved in one batch, then the reference's first continuation (KD -> 310,
// tests/testGoddard.cpp:105) on the whole batch.  Problem 0 is the unperturbed reference problem and is
// also solved alone through `shooting` to show that batch membership does not change a result.
//   usage: demo_goddard_batch [B=256] [numDevice=1]      (numDevice 0: every visible GPU shares the batch)
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "socp/shooting.hpp"
#include "socp/shooting_batch.hpp"
#include "models/goddard/goddard.hpp"

static double unit(unsigned long long & s) {		// uniform in [-1, 1]
	s = s * 6364136223846793005ULL + 1442695040888963407ULL;
	return ((double)(s >> 11) / 9007199254740992.0) * 2.0 - 1.0;
}

int main(int argc, char **argv) {
	const long B = argc > 1 ? atol(argv[1]) : 256;
	const int numDevice = argc > 2 ? atoi(argv[2]) : 1;
	const int M = 6;
	goddard my_goddard("", 10);
	const int dim = my_goddard.GetDim();
	std::vector<int> mode_Xf(dim, model::FREE);
	mode_Xf[0] = mode_Xf[1] = mode_Xf[2] = model::FIXED;

	// reference problem data (tests/testGoddard.cpp:53-76)
	model::mstate Xi0(2 * dim, 0.1), Xf0(2 * dim, 0.0);
	Xi0[0] = 0.999949994; Xi0[1] = 0.0001; Xi0[2] = 0.01; Xi0[3] = 1e-10; Xi0[4] = 1e-10; Xi0[5] = 1e-10; Xi0[6] = 1.0;
	Xf0[0] = 1.01;
	const real ti = 0, tf = 0.1;

	// the guess for the interior nodes is integrated with the constructor's KD = 310; the solve uses KD = 0
	my_goddard.SetParameterDataName("mu2", 1.0);
	shooting_batch batch(my_goddard, M, B, numDevice);
	printf("batch of %ld problems on %d GPU(s)\n", B, batch.GetNumDevice());
	batch.SetPrecision(1e-6);
	batch.SetMode(model::FREE, mode_Xf);
	std::vector<real> vti(B, ti), vtf(B, tf);
	std::vector<model::mstate> vXi(B, Xi0), vXf(B, Xf0);
	unsigned long long seed = 20260002ULL;
	for (long k = 1; k < B; k++) {
		for (int j = 0; j < 3; j++) vXi[k][j] *= 1 + 1e-3 * unit(seed);
		vXi[k][6] *= 1 + 1e-2 * unit(seed);
		vXf[k][0] = 1.01 + 0.002 * unit(seed);
	}
	batch.InitShooting(vti, vXi, vtf, vXf);
	for (long k = 0; k < B; k++) {
		std::vector<real> block = batch.GetModelParameters(k);
		block[2] = 0.0;									// KD = 0
		batch.SetModelParameters(k, block);
	}
	long ok = batch.SolveOCP(0.0);
	printf("batch %ld solved: %ld converged (info == 1)\n", B, ok);
	printf("problem 0: info %d nfev %d tf %.17g p0", batch.GetInfo(0), batch.GetCallNumber(0)[0], batch.GetParameters(0, 2 * dim * M));
	for (int i = 0; i < dim; i++) printf(" %.17g", batch.GetParameters(0, dim + i));
	printf("\n");

	// the same problem alone, through the single-problem class
	shooting single(my_goddard, M, 1);
	single.SetPrecision(1e-6);
	single.SetMode(model::FREE, mode_Xf);
	single.InitShooting(ti, Xi0, tf, Xf0);
	my_goddard.SetParameterDataName("KD", 0.0);
	const int info1 = single.SolveOCP(0.0);
	printf("single   : info %d nfev %d tf %.17g p0", info1, single.GetCallNumber()[0], single.GetParameters(2 * dim * M));
	for (int i = 0; i < dim; i++) printf(" %.17g", single.GetParameters(dim + i));
	printf("\n");
	bool same = (info1 == batch.GetInfo(0));
	for (int i = 0; i < 2 * dim * M + 1; i++) same = same && (single.GetParameters(i) == batch.GetParameters(0, i));
	printf("batch member 0 identical to the single solve: %s\n", same ? "yes" : "NO");

	// continuation KD -> 310 on every converged problem of the batch
	std::vector<real> goal(B, 310.0);
	long ok2 = batch.SolveOCP(1.0, 2, goal);
	printf("continuation KD -> 310: %ld problems arrived (info == 1)\n", ok2);
	return same ? 0 : 1;
}
