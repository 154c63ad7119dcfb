// cpp/src/models/goddard/goddard.hpp -- mirror of src/models/goddard/goddard.hpp:19-112.
// Generalised Goddard problem (Bonnans, Martinon, Trelat, JOTA 2008); RHS, control, singular arc
// and Hamiltonian run on the device (socp_b200/csrc/models.cuh, Model<GODDARD>).
#include "../../socp/model.hpp"
#include "../../socp/map.hpp"

#include <iostream>

#ifndef _GODDARD_H_
#define _GODDARD_H_

class goddard:public model
{
public:
	struct parameters_struct{
		real C = 3.5;
		real b = 7.0;
		real KD = 310.0;
		real kr = 500.0;
		real u_max = 1.0;
		real mu1 = 1.0;
		real mu2 = 0.0;
		real singularControl = -1;
	};

	goddard(std::string the_fileTrace = std::string(""), int stepNbr = 10);
	virtual ~goddard();

	virtual mstate SwitchingTimesFunction(real const& t, mstate const& X, mstate const& Xp, int isJac) const;
	void SetParameterDataName(std::string name, real value);
	real & GetParameterDataName(std::string name);

	virtual int DeviceModelId() const;
	virtual std::vector<real> DeviceParams() const;
	virtual void TraceTail(real H, real extra, std::ostream & file) const;
};

#endif //_GODDARD_H_
