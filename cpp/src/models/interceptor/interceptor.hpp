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
		real c0;				///< max curvature at ground level (1/m)
		real hr;				///< reference altitude (m)
		real d0;				///< drag at ground level (1/m)
		real eta;				///< coeff of efficiency
		real propellant_mass;	///< propellant mass (kg)
		real empty_mass; 		///< empty mass (kg)
		real q; 				///< mass flow rate (kg/s)
		real ve; 				///< gas speed (m/s)
		real alpha_max; 		///< max angle of attack (rd)
		real u_max;				///< max of the normalized control (for saturation)
		real a_max;				///< max acceleration allowed (for saturation)
		real r_2p;				///< ratio of used gas for the second propulsion phase
		real t_2p;				///< start time for the second propulsion phase
		real mu_gft;			///< parameter for considering gravity and propulsion
		real muT;				///< weight for time cost
		real muV;				///< weight for velocity cost
		real muC;				///< weight for quadratic control cost
	};

	interceptor(std::string the_fileTrace = std::string(""));
	virtual ~interceptor();
	parameters_struct & GetParameterData();
	void InitAnalytical(real const& ti, mstate & Xi, real const& tf, mstate & Xf) const;

	virtual int DeviceModelId() const;
	virtual std::vector<real> DeviceParams() const;
	virtual int DeviceSteps() const { return 50; }		///< interceptor.cpp:54 (per stage)

private:
	struct data_struct;
	data_struct *data;
};

#endif //_INTERCEPTOR_H_
