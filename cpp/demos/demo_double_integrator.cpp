// cpp/demos/demo_double_integrator.cpp -- the scenario of the reference's double-integrator demo
// (tests/testDoubleIntegrator.cpp:25-143) restated against the mirror API: a free-final-time solve,
// a continuation on the boundary data (Xf[1] -> 20) and a continuation on a model parameter
// (muT -> 0.02).  Prints one machine-readable line per stage:
//   stage <k> info <i> nfev <n> tf <%.17g> p0 <six costates %.17g>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "socp/shooting.hpp"
#include "models/doubleIntegrator/doubleIntegrator.hpp"

static void report(int stage, int info, shooting & s, int dim, int numParam) {
	std::vector<int> calls = s.GetCallNumber();
	printf("stage %d info %d nfev %d njev %d tf %.17g p0", stage, info, calls[0], calls[1], s.GetParameters(numParam - 1));
	for (int i = 0; i < dim; i++) printf(" %.17g", s.GetParameters(dim + i));
	printf("\n");
}

int main(int argc, char **argv) {
	const int order = argc > 1 ? atoi(argv[1]) : 0;	// 0: forward-difference Jacobian (hybrd); 1: analytic (hybrj)
	doubleIntegrator my_model(order, "");
	const int dim = my_model.GetDim();
	shooting my_shooting(my_model, 1, 1);
	my_shooting.SetPrecision(1e-8);
	std::vector<int> mode_X(dim, model::FIXED);
	my_shooting.SetMode(model::FREE, mode_X);
	const int numParam = 2 * dim + 1;

	model::mstate Xi(2 * dim), Xf(2 * dim);
	const real ti = 0, tf = 10;
	for (int i = 0; i < dim; i++) { Xi[i] = 0; Xi[dim + i] = 0.01; Xf[i] = 0; Xf[dim + i] = 0; }
	Xf[0] = 10; Xf[1] = 15;
	my_shooting.InitShooting(ti, Xi, tf, Xf);
	int info = my_shooting.SolveOCP(0.0);
	report(1, info, my_shooting, dim, numParam);

	Xf[1] = 20;
	my_shooting.SetDesiredState(ti, Xi, tf, Xf);
	info = my_shooting.SolveOCP(1.0);
	report(2, info, my_shooting, dim, numParam);

	info = my_shooting.SolveOCP(1.0, my_model.GetParameterData().muT, 0.02);
	report(3, info, my_shooting, dim, numParam);

	model::mstate X_end = my_shooting.Move(1e30);		// clamped to the final time
	printf("final state");
	for (int i = 0; i < dim; i++) printf(" %.12g", X_end[i]);
	printf("\n");
	return info == 1 ? 0 : 1;
}
