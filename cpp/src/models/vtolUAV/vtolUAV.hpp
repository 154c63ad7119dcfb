// cpp/src/models/vtolUAV/vtolUAV.hpp -- mirror of src/models/vtolUAV/vtolUAV.hpp:17-113.
// The map must be an `obstacle` (src/maps/obstacle): its table is uploaded to the device.
#include "../../socp/model.hpp"
#include "../../socp/map.hpp"

#include <iostream>

#ifndef _VTOLUAV_H_
#define _VTOLUAV_H_

class vtolUAV:public model
{
public:
	struct parameters_struct{
		real u_max;
		real a_max;
		real alphaT;
		real alphaV;
		real invSigmaXwp;
		real Vd;
		real ca;
		int nWP_tot;
		int nWP;
	};

	vtolUAV(map & the_map, std::string the_fileTrace = std::string(""));
	virtual ~vtolUAV();
	parameters_struct & GetParameterData();
	map & GetMap() const;

	virtual int DeviceModelId() const;
	virtual std::vector<real> DeviceParams() const;
	virtual int DeviceSteps() const { return 100; }		///< vtolUAV.cpp:38

private:
	map & myMap;
	struct data_struct;
	data_struct *data;
};

#endif //_VTOLUAV_H_
