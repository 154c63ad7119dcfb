// cpp/src/models/doubleIntegrator/doubleIntegrator.cpp -- host side of the double-integrator mirror
// (reference: src/models/doubleIntegrator/doubleIntegrator.cpp); dynamics on the device
// (Model<DOUBLE_INTEGRATOR>).
#include "doubleIntegrator.hpp"
#include "../../../../include/socp_b200.h"

struct doubleIntegrator::data_struct {
	parameters_struct parameters;
	int requestedSteps;			// what SetStepNumber stores; the integration keeps 30 steps as in the reference
};

doubleIntegrator::doubleIntegrator(int modelOrder, std::string the_fileTrace) : model(6, modelOrder, 30, the_fileTrace) {
	data = new data_struct;
	data->parameters.u_max = 1;
	data->parameters.a_max = 1;
	data->parameters.muT = 0.01;
	data->requestedSteps = 30;
}

doubleIntegrator::~doubleIntegrator() { delete data; }

doubleIntegrator::parameters_struct & doubleIntegrator::GetParameterData() { return data->parameters; }

void doubleIntegrator::SetStepNumber(int step) { data->requestedSteps = step; }

int doubleIntegrator::DeviceModelId() const { return SOCP_DOUBLE_INTEGRATOR; }

std::vector<real> doubleIntegrator::DeviceParams() const {
	const parameters_struct & p = data->parameters;
	const real block[3] = {p.u_max, p.a_max, p.muT};
	return std::vector<real>(block, block + 3);
}
