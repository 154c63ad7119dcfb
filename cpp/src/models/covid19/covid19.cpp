// cpp/src/models/covid19/covid19.cpp -- host side of the SEIR mirror
// (reference: src/models/covid19/covid19.cpp); dynamics on the device (Model<COVID19>).
#include "covid19.hpp"
#include "../../../../include/socp_b200.h"

struct covid19::data_struct {
	parameters_struct parameters;
};

covid19::covid19(std::string the_fileTrace) : model(4, 0, 1000, the_fileTrace) {
	data = new data_struct;
	parameters_struct & p = data->parameters;			// covid19.cpp:29-36
	p.R0 = 4; p.Tinf = 10; p.Tinc = 5; p.N = 1; p.Imax = 0.1; p.muI = 1; p.umin = -10; p.umax = 20;
}

covid19::~covid19() { delete data; }

covid19::parameters_struct & covid19::GetParameterData() { return data->parameters; }

int covid19::DeviceModelId() const { return SOCP_COVID19; }

std::vector<real> covid19::DeviceParams() const {
	const parameters_struct & p = data->parameters;
	const real block[8] = {p.R0, p.Tinf, p.Tinc, p.N, p.Imax, p.muI, p.umin, p.umax};
	return std::vector<real>(block, block + 8);
}
