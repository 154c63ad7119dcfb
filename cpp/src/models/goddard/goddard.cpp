// cpp/src/models/goddard/goddard.cpp -- host side of the goddard mirror
// (reference: src/models/goddard/goddard.cpp).  The dynamics, control law, singular arc and
// Hamiltonian live on the device (socp_b200/csrc/models.cuh, Model<GODDARD>); this file only keeps
// the parameter map the reference exposes for continuation and packs it for the C ABI.
#include "goddard.hpp"
#include "../../../../include/socp_b200.h"

goddard::goddard(std::string the_fileTrace, int stepNbr) : model(7, 0, stepNbr, the_fileTrace) {
	const parameters_struct d;							// constructor defaults (goddard.hpp:29-36)
	parameters["C"] = d.C;
	parameters["b"] = d.b;
	parameters["KD"] = d.KD;
	parameters["kr"] = d.kr;
	parameters["u_max"] = d.u_max;
	parameters["mu1"] = d.mu1;
	parameters["mu2"] = d.mu2;
	parameters["singularControl"] = d.singularControl;
}

goddard::~goddard() {}

int goddard::DeviceModelId() const { return SOCP_GODDARD; }

std::vector<real> goddard::DeviceParams() const {
	// block order of include/socp_b200.h; an unknown key throws std::out_of_range as parameters.at does
	// in the reference (goddard.cpp:71)
	static const char *names[8] = {"C", "b", "KD", "kr", "u_max", "mu1", "mu2", "singularControl"};
	std::vector<real> block(8);
	for (int k = 0; k < 8; k++) block[k] = parameters.at(names[k]);
	return block;
}

// with FREE interior times the reference uses H itself as switching condition (goddard.cpp:343-370)
model::mstate goddard::SwitchingTimesFunction(real const& t, mstate const& X, mstate const&, int isJac) const {
	return Hamiltonian(t, X, isJac);
}

void goddard::SetParameterDataName(std::string name, real value) { parameters.at(name) = value; }

real & goddard::GetParameterDataName(std::string name) { return parameters.at(name); }

// goddard::Trace appends the switching function after H (goddard.cpp:337-339)
void goddard::TraceTail(real H, real extra, std::ostream & file) const { file << H << "\t" << extra << std::endl; }
