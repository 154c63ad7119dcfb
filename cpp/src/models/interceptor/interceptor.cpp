// cpp/src/models/interceptor/interceptor.cpp -- host side of the interceptor mirror
// (reference: src/models/interceptor/interceptor.cpp).  The two charts, two stages, chart conversion
// and the dynamics run on the device (Model<INTERCEPTOR>).  What stays on the host is what the
// reference also runs once per problem before any solve: the closed-form costate guess
// (InitAnalytical, interceptor.cpp:844-955) and the mass bookkeeping it needs (ComputeMass, :984-999).
#include <cmath>

#include "interceptor.hpp"
#include "../../../../include/socp_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct interceptor::data_struct {
	parameters_struct parameters;
	real R_Earth, mu0;
	int stageMode;				// 1 = powered, 0 = coasting: what the last ComputeTraj left behind (interceptor.cpp:29)
};

interceptor::interceptor(std::string the_fileTrace) : model(6, 0, 50, the_fileTrace) {
	data = new data_struct;
	parameters_struct & p = data->parameters;			// interceptor.cpp:36-50
	p.c0 = 0.00075; p.hr = 7500; p.d0 = 0.00005; p.eta = 0.442;
	p.propellant_mass = 200; p.empty_mass = 200; p.q = 10; p.ve = 1500;
	p.alpha_max = M_PI / 6; p.u_max = 1; p.a_max = 1500;
	p.r_2p = 0; p.t_2p = 0;
	p.mu_gft = 1; p.muT = 0; p.muV = 1; p.muC = 0;
	data->R_Earth = 6378145;
	data->mu0 = 3.986e14;
	data->stageMode = 0;
}

interceptor::~interceptor() { delete data; }

interceptor::parameters_struct & interceptor::GetParameterData() { return data->parameters; }

int interceptor::DeviceModelId() const { return SOCP_INTERCEPTOR; }

std::vector<real> interceptor::DeviceParams() const {
	const parameters_struct & p = data->parameters;
	const real block[17] = {p.c0, p.hr, p.d0, p.eta, p.propellant_mass, p.empty_mass, p.q, p.ve, p.alpha_max,
	                        p.u_max, p.a_max, p.r_2p, p.t_2p, p.mu_gft, p.muT, p.muV, p.muC};
	return std::vector<real>(block, block + 17);
}

int interceptor::GetMode(real const&, mstate const&) const { return data->stageMode; }

// the staged integration itself is one device call; the stage flag is left as interceptor.cpp:165-220
// leaves it (powered iff the segment starts and ends before burn-out)
model::mstate interceptor::ComputeTraj(real const& t0, mstate const& X0, real const& tf, int isTrace, int isJac) {
	const real t1 = data->parameters.propellant_mass / data->parameters.q;
	data->stageMode = (t0 < t1 && !(tf > t1)) ? 1 : 0;
	return model::ComputeTraj(t0, X0, tf, isTrace, isJac);
}

void interceptor::TraceTail(real H, real extra, std::ostream & file) const { file << H << "\t" << (int)extra << std::endl; }

real interceptor::ComputeMass(real const& t, mstate const& X) const {
	const parameters_struct & p = data->parameters;
	const real burnt = p.q * p.mu_gft * (GetMode(t, X) == 1 ? t : p.propellant_mass / p.q);
	return p.empty_mass + p.propellant_mass - burnt;
}

namespace {
struct V3 { real x, y, z; };
inline V3 operator+(V3 a, V3 b) { V3 c = {a.x + b.x, a.y + b.y, a.z + b.z}; return c; }
inline V3 operator-(V3 a, V3 b) { V3 c = {a.x - b.x, a.y - b.y, a.z - b.z}; return c; }
inline V3 operator*(real s, V3 a) { V3 c = {s * a.x, s * a.y, s * a.z}; return c; }
inline real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 up(real L, real l) { V3 c = {cos(L) * cos(l), cos(L) * sin(l), sin(L)}; return c; }			// local vertical
inline V3 north(real L, real l) { V3 c = {-sin(L) * cos(l), -sin(L) * sin(l), cos(L)}; return c; }
inline V3 east(real l) { V3 c = {-sin(l), cos(l), 0}; return c; }
}

// Closed-form costate guess from the proportional-navigation-like mid-course law of the IFAC 2017
// paper: gains k1, k2, k3 of the exponential-atmosphere guidance, line-of-sight angles lambda1 (vertical
// plane) and lambda2 (horizontal plane), the guidance commands u1, u2 and their time derivatives, then
// the costates that make those commands the PMP minimisers.  Only Xi[6..12) is written.
void interceptor::InitAnalytical(real const& ti, mstate & Xi, real const&, mstate & Xf) const {
	const parameters_struct & p = data->parameters;
	const real h = Xi[0], v = Xi[1], gam = Xi[2], chi = Xi[3], L = Xi[4], l = Xi[5];
	const real hf = Xf[0], gamf = Xf[2], chif = Xf[3], Lf = Xf[4], lf = Xf[5];
	const real eta = p.eta, hr = p.hr;
	const real massRatio = (p.propellant_mass + p.empty_mass) / ComputeMass(ti, Xi);
	const real C = p.c0 * exp(-h / hr) * massRatio;			// max curvature at altitude h
	const real d = p.d0 * exp(-h / hr) * massRatio;			// drag at altitude h
	const real r = h + data->R_Earth, rf = hf + data->R_Earth;
	const real sg = sin(gam), cg = cos(gam), tg = tan(gam);

	// geometry: positions in the Earth frame, line of sight D, range R and range rate
	const V3 e = up(L, l), ef = up(Lf, lf), n = north(L, l), ea = east(l);
	const V3 D = rf * ef - r * e;
	const real R = sqrt(dot(D, D));
	const V3 w = sg * e + (cg * cos(chi)) * n + (cg * sin(chi)) * ea;	// unit velocity
	const real Rdot = -dot(D, w) / R;

	// guidance gains and their derivatives along the trajectory (s = b R)
	const real b = sqrt(C * d / (2 * eta));
	const real bdot = -C * d * sg * sqrt(2 * eta / (C * d)) / (2 * eta * hr);
	const real s = b * R, sdot = bdot * R + b * Rdot, Ep = exp(s), Em = exp(-s);
	const real den = 4 + Ep * (s - 2) - Em * (s + 2);
	const real dden = sdot * (Ep * (s - 2) + Em * (s + 2) + Ep - Em);
	const real N1 = Ep - Em - 2 * s, dN1 = sdot * (Ep + Em - 2);
	const real N2 = Ep * (s - 1) + Em * (s + 1), dN2 = s * sdot * (Ep - Em);
	const real k1 = s * N1 / den, k1dot = sdot * N1 / den + s * (dN1 * den - dden * N1) / (den * den);
	const real k2 = s * N2 / den, k2dot = sdot * N2 / den + s * (dN2 * den - dden * N2) / (den * den);
	const real k3 = 2 + k1 - k2, k3dot = k1dot - k2dot;

	// lambda1: elevation of the line of sight (sign from which side of the local horizontal the target is)
	real lam1;
	{
		const real dist = fabs(rf * dot(e, ef) - r), below = -dot(e, D);
		if (R == 0) lam1 = gamf;
		else if (dist / R >= 1) lam1 = (below > 0) ? -M_PI / 2.0 : M_PI / 2.0;
		else lam1 = (below > 0) ? -asin(dist / R) : asin(dist / R);
	}
	// lambda2: azimuth of the target position projected on the local horizontal plane
	real lam2;
	{
		const V3 proj = rf * ef - (rf * dot(ef, e)) * e;
		const real len = sqrt(dot(proj, proj)), c = dot(proj, n);
		if (len == 0) lam2 = 0;
		else if (c / len <= -1) lam2 = M_PI;
		else if (c / len >= 1) lam2 = 0;
		else lam2 = (dot(proj, ea) >= 0) ? acos(c / len) : -acos(c / len);
	}

	// guidance commands and their derivatives
	const real e1 = gam - lam1, e2 = chi - lam2, R2 = R * R;
	const real u1 = -(k1 * (gamf - lam1) / R + k2 * sin(e1) / R + k3 * cg / (2 * hr)) / C;
	const real u2 = -(k1 * (chif - lam2) * cg / R + k2 * sin(e2) * cg / R) / C;
	const real carry = sg / (C * hr);						// (dC/dt)/C^2 along the path
	const real du1 = carry * C * u1
		- (k1dot * (gamf - lam1) / R + k1 * sin(e1) / R2 - k1 * (gamf - lam1) * Rdot / R2 + k2dot * sin(e1) / R
		   + k2 * cos(e1) * (C * u1 + sin(e1) / R) / R - k2 * sin(e1) * Rdot / R2 + k3dot * cg / (2 * hr)
		   - C * u1 * k3 * sg / (2 * hr)) / C;
	const real du2 = carry * C * u2
		- (k1dot * cg * (chif - lam2) / R - k1 * C * u1 * sg * (chif - lam2) / R + k1 * cg * sin(e2) / R2
		   - k1 * cg * (chif - lam2) * Rdot / R2 + k2dot * cg * sin(e2) / R - k2 * sg * sin(e2) * C * u1 / R
		   + k2 * cg * cos(e2) * (C * u2 / cg + sin(e2) / R) / R - k2 * cg * sin(e2) * Rdot / R2) / C;

	// costates (per unit speed first): p_v = -1, p_gamma and p_chi from the minimising controls,
	// p_h, p_L, p_l from the stationarity of H along the guidance law
	const real pg = 2 * eta * u1, pc = 2 * eta * u2 * cg;
	const real A = C * u1 * pg + eta * C * (u1 * u1 + u2 * u2) + d;
	const real Wm = C * u1 * u2 * tg - du2;
	const real B = C * u2 * pc * (cg - sg * tg) + 2 * eta * sg * cg * du1 + cg * cg * A;
	const real ph = (-sg * cg * A - 2 * sg * C * u2 * pc + 2 * eta * cg * cg * du1) / cg;
	const real pL = r * (-cos(chi) * B + 2 * eta * cg * sin(chi) * Wm) / cg;
	const real pl = -r * cos(L) * (sin(chi) * B + 2 * eta * cg * cos(chi) * Wm) / cg;
	Xi[6] = v * ph;
	Xi[7] = -1;
	Xi[8] = v * pg;
	Xi[9] = v * pc;
	Xi[10] = v * pL;
	Xi[11] = v * pl;
}
