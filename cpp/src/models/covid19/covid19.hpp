// cpp/src/models/covid19/covid19.hpp -- mirror of src/models/covid19/covid19.hpp:21-84 (SEIR model).
#include "../../socp/model.hpp"
#include "../../socp/map.hpp"

#include <iostream>

#ifndef _COVID19_H_
#define _COVID19_H_

class covid19:public model
{
public:
	struct parameters_struct{
		real R0;
		real Tinf;
		real Tinc;
		real N;
		real Imax;
		real muI;
		real umin;
		real umax;
	};

	covid19(std::string the_fileTrace = std::string(""));
	virtual ~covid19();
	parameters_struct & GetParameterData();

	virtual int DeviceModelId() const;
	virtual std::vector<real> DeviceParams() const;
	virtual int DeviceSteps() const { return 1000; }		///< covid19.cpp:38

private:
	struct data_struct;
	data_struct *data;
};

#endif //_COVID19_H_
