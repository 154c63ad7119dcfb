// cpp/src/models/vtolUAV/vtolUAV.cpp -- host side of the VTOL UAV mirror
// (reference: src/models/vtolUAV/vtolUAV.cpp); dynamics and the obstacle penalty field on the device
// (Model<VTOL_UAV>, obstacle_eval).  The map must be an `obstacle`: its table is uploaded to the
// engine context whenever the parameter block is packed.
#include <cstdlib>

#include "vtolUAV.hpp"
#include "../../maps/obstacle/obstacle.hpp"
#include "../../../../include/socp_b200.h"

struct vtolUAV::data_struct {
	parameters_struct parameters;
};

vtolUAV::vtolUAV(map & the_map, std::string the_fileTrace) : model(6, 0, 100, the_fileTrace), myMap(the_map) {
	data = new data_struct;
	parameters_struct & p = data->parameters;			// vtolUAV.cpp:27-35
	p.u_max = 10; p.a_max = 0.3; p.alphaT = 0.05; p.alphaV = 0; p.invSigmaXwp = 1. / 60; p.Vd = 1; p.ca = 0;
	p.nWP_tot = 0; p.nWP = 0;
}

vtolUAV::~vtolUAV() { delete data; }

vtolUAV::parameters_struct & vtolUAV::GetParameterData() { return data->parameters; }

map & vtolUAV::GetMap() const { return myMap; }

int vtolUAV::DeviceModelId() const { return SOCP_VTOL_UAV; }

std::vector<real> vtolUAV::DeviceParams() const {
	obstacle *obs = dynamic_cast<obstacle *>(&myMap);
	if (!obs) {
		std::cerr << std::endl << "ERROR : the B200 engine evaluates the vtolUAV map on the device and needs an `obstacle` map" << std::endl;
		exit(1);
	}
	obs->Upload();
	const parameters_struct & p = data->parameters;
	const obstacle::parameters_struct & o = obs->GetParameterData();
	const real block[13] = {p.u_max, p.a_max, p.alphaT, p.alphaV, p.invSigmaXwp, p.Vd, p.ca, (real)p.nWP_tot, (real)p.nWP,
	                        o.phiObs, o.psiWP, o.muObs, o.sigmaWP};
	return std::vector<real>(block, block + 13);
}
