// cpp/src/models/interceptor/interceptor.hpp -- mirror of
// src/models/interceptor/interceptor.hpp:21-143 (Bonalli, Herisse, Trelat, IFAC WC 2017).
// Two charts, two stages and the chart conversion run on the device (Model<INTERCEPTOR>);
// InitAnalytical (the closed-form costate guess, interceptor.cpp:844-955) is host glue here as in
// the reference.
#include "../../socp/model.hpp"
#include "../../socp/map.hpp"

#include <iostream>

#ifndef _INTERCEPTOR_H_
#define _INTERCEPTOR_H_

class interceptor:public model
{
public:
	struct parameters_struct{
		real c0;
		real hr;
		real d0;
		real eta;
		real propellant_mass;
		real empty_mass;
		real q;
		real ve;
		real alpha_max;
		real u_max;
		real a_max;
		real r_2p;
		real t_2p;
		real mu_gft;
		real muT;
		real muV;
		real muC;
	};

	interceptor(std::string the_fileTrace = std::string(""));
	virtual ~interceptor();
	parameters_struct & GetParameterData();
	void InitAnalytical(real const& ti, mstate & Xi, real const& tf, mstate & Xf) const;

	virtual mstate ComputeTraj(real const& t0, mstate const& X0, real const& tf, int isTrace, int isJac);
	virtual int GetMode(real const& t, mstate const& X) const;

	virtual int DeviceModelId() const;
	virtual std::vector<real> DeviceParams() const;
	virtual int DeviceSteps() const { return 50; }		///< interceptor.cpp:54 (per stage)
	virtual void TraceTail(real H, real extra, std::ostream & file) const;

private:
	real ComputeMass(real const& t, mstate const& X) const;
	struct data_struct;
	data_struct *data;
};

#endif //_INTERCEPTOR_H_
