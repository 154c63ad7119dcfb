// cpp/src/socp/map.hpp -- mirror of src/socp/map.hpp:15-47 (abstract penalty field).
#include <vector>
#include "commonType.hpp"

#ifndef _MAP_H_
#define _MAP_H_

class map
{
public:
	map() {};
	virtual ~map() {};
	virtual void Function(std::vector<real> const& state, real & func) const = 0;
	virtual void Gradient(std::vector<real> const& state, std::vector<real> & grad) const = 0;
};

#endif //_MAP_H_
