// cpp/src/maps/obstacle/obstacle.hpp -- mirror of src/maps/obstacle/obstacle.hpp:12-97: box /
// ellipsoid obstacle table read from text files; the tanh penalty and its gradient are evaluated
// on the device inside the vtolUAV RHS (socp_b200/csrc/models.cuh, obstacle_eval).
#include <string>
#include "../../socp/map.hpp"

#ifndef _OBSTACLE_H_
#define _OBSTACLE_H_

class obstacle:public map
{
public:
	struct parameters_struct{
		real phiObs;
		real psiWP;
		real muObs;
		real sigmaWP;
	};

	obstacle(std::string the_fileObstacles = std::string(""), std::string the_fileWP = std::string(""));
	virtual ~obstacle();

	/// host entry points of the abstract map: evaluated on the device (a vtolUAV point evaluation)
	virtual void Function(std::vector<real> const& position, real & funcTot) const;
	virtual void Gradient(std::vector<real> const& position, std::vector<real> & gradTot) const;

	parameters_struct & GetParameterData();
	const std::vector<std::vector<real>> & GetPath();

	// ---- B200 engine hooks ------------------------------------------------------------------------
	int Count() const;
	/// obstacle table as {type[n], pos[n][3], rad[n][3]} for socp_set_obstacles
	void Table(std::vector<real> & type, std::vector<real> & pos, std::vector<real> & rad) const;
	/// push the table to the engine context (socp_set_obstacles); done lazily, once
	void Upload() const;

private:
	struct data_struct;
	data_struct *data;
	void ReadObstacleInput();
	void ReadWPInput();
};

#endif //_OBSTACLE_H_
