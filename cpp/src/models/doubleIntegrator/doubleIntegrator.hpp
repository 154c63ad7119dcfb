// cpp/src/models/doubleIntegrator/doubleIntegrator.hpp -- mirror of
// src/models/doubleIntegrator/doubleIntegrator.hpp:16-79.  modelOrder 0: forward-difference Powell
// hybrid (hybrd); modelOrder 1: analytic Jacobian from the variational integration (hybrj), both on
// the device.
#include "../../socp/model.hpp"

#include <iostream>

#ifndef _DOUBLEINTEGRATOR_H_
#define _DOUBLEINTEGRATOR_H_

class doubleIntegrator : public model
{
public:
	struct parameters_struct{
		real u_max;
		real a_max;
		real muT;
	};

	doubleIntegrator(int modelOrder, std::string the_fileTrace);
	virtual ~doubleIntegrator();

	parameters_struct & GetParameterData();
	void SetStepNumber(int step);		///< writes an unused field in the reference (doubleIntegrator.cpp:313): 30 steps stay

	virtual int DeviceModelId() const;
	virtual std::vector<real> DeviceParams() const;

private:
	struct data_struct;
	data_struct *data;
};

#endif //_DOUBLEINTEGRATOR_H_
