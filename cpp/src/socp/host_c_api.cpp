// cpp/src/socp/host_c_api.cpp -- a few extern "C" doors into the C++ mirror for the Python tests
// (host-only pieces that need no GPU can be checked against the reference on the CPU).
#include "../models/interceptor/interceptor.hpp"

extern "C" {

// interceptor::InitAnalytical (interceptor.cpp:844-955) with the given parameter block
// (layout of include/socp_b200.h); Xi[6..12) receives the costate guess.
int socp_host_interceptor_init_analytical(const double *params, double ti, double *Xi, double tf, double *Xf) {
	interceptor m("");
	interceptor::parameters_struct & p = m.GetParameterData();
	p.c0 = params[0]; p.hr = params[1]; p.d0 = params[2]; p.eta = params[3]; p.propellant_mass = params[4];
	p.empty_mass = params[5]; p.q = params[6]; p.ve = params[7]; p.alpha_max = params[8]; p.u_max = params[9];
	p.a_max = params[10]; p.r_2p = params[11]; p.t_2p = params[12]; p.mu_gft = params[13]; p.muT = params[14];
	p.muV = params[15]; p.muC = params[16];
	model::mstate a(Xi, Xi + 12), b(Xf, Xf + 12);
	m.InitAnalytical(ti, a, tf, b);
	for (int i = 0; i < 12; i++) Xi[i] = a[i];
	return 0;
}

}
