// cpp/src/socp/odeTools.hpp -- mirror of the reference's odeTools interface
// (src/socp/odeTools.hpp:17-191).  The integrator itself (RK4, odeTools.cpp:89-146) runs on the
// GPU behind model::ComputeTraj; the host-side function-pointer helpers of the reference
// (RK1/RK2/RK4 over a user callback) cannot run on a device and are not provided.
#include "commonType.hpp"

#include <sstream>
#include <vector>

#ifndef _ODETOOLS_H_
#define _ODETOOLS_H_

class odeTools
{
public:
	typedef std::vector<real> odeVector;

	odeTools() : odeIntTol(1e-8) {};
	virtual ~odeTools() {};

	real odeIntTol;								///< precision if dopri5 is used

	virtual void SetODEIntPrecision(real const& xtol) { odeIntTol = xtol; };

	/// dX/dt = Model(t, X): evaluated on the device for the shipped models
	virtual odeVector Model(real const& t, odeVector const& X, int isJac = 0) const = 0;

	virtual void Trace(real const& t, odeVector const& X, std::stringstream & file) const = 0;

	static odeVector MultState(real a, odeVector const& X) {
		odeVector Y(X.size());
		for (size_t i = 0; i < X.size(); i++) Y[i] = a*X[i];
		return Y;
	};
	static odeVector AddState(odeVector const& X, odeVector const& Y) {
		odeVector Z(X.size());
		for (size_t i = 0; i < X.size(); i++) Z[i] = X[i] + Y[i];
		return Z;
	};
};

#endif //_ODETOOLS_H_
