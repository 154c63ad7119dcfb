// cpp/src/maps/obstacle/obstacle.cpp -- host side of the obstacle map mirror
// (reference: src/maps/obstacle/obstacle.cpp).  Parses the same text files; the tanh penalty and its
// gradient (obstacle.cpp:183-316) are evaluated on the device inside the vtolUAV dynamics.
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "obstacle.hpp"
#include "../../socp/model.hpp"
#include "../../../../include/socp_b200.h"

struct obstacle::data_struct {
	std::vector<real> type;								// 0 ellipsoid, 1 box
	std::vector< std::vector<real> > position, radius;	// [n][3]
	std::vector< std::vector<real> > waypoints;			// [n_wp][6]: position, unit direction to the next one
	parameters_struct parameters;
	std::string fileObstacles, fileWP;
	bool uploaded;
};

// one "label line, then rows" block of the input format
static bool read_rows(std::ifstream & is, int rows, int cols, std::vector< std::vector<real> > & out) {
	std::string line;
	std::getline(is, line);								// label
	out.assign(rows, std::vector<real>(cols, 0));
	for (int i = 0; i < rows; i++) {
		if (!std::getline(is, line)) return false;
		std::istringstream row(line);
		for (int k = 0; k < cols; k++)
			if (!(row >> out[i][k])) return false;
	}
	return true;
}

static int read_count(std::ifstream & is) {
	std::string line;
	int n = 0;
	std::getline(is, line);								// label
	std::getline(is, line);
	std::istringstream s(line);
	return (s >> n) ? n : -1;
}

obstacle::obstacle(std::string the_fileObstacles, std::string the_fileWP) {
	data = new data_struct;
	data->parameters.phiObs = 1;						// obstacle.cpp:47-50
	data->parameters.psiWP = 0.03;
	data->parameters.muObs = 1;
	data->parameters.sigmaWP = 2.5;
	data->fileObstacles = the_fileObstacles;
	data->fileWP = the_fileWP;
	data->uploaded = false;
	ReadObstacleInput();
	ReadWPInput();
}

obstacle::~obstacle() { delete data; }

// file layout: "n:" count, "type:" n rows, "position:" n rows of 3, "radius:" n rows of 3
void obstacle::ReadObstacleInput() {
	std::ifstream is(data->fileObstacles.c_str());
	if (!is) return;
	const int n = read_count(is);
	if (n < 0) return;
	std::vector< std::vector<real> > types;
	if (!read_rows(is, n, 1, types)) return;
	data->type.resize(n);
	for (int i = 0; i < n; i++) data->type[i] = types[i][0];
	if (!read_rows(is, n, 3, data->position)) return;
	read_rows(is, n, 3, data->radius);
}

// file layout: "n_wp:" count, "position_wp:" rows of 3; directions point to the next waypoint
void obstacle::ReadWPInput() {
	std::ifstream is(data->fileWP.c_str());
	if (!is) return;
	const int n = read_count(is);
	if (n < 0) return;
	std::vector< std::vector<real> > pos;
	if (!read_rows(is, n, 3, pos)) return;
	data->waypoints.assign(n, std::vector<real>(6, 0));
	for (int i = 0; i < n; i++)
		for (int k = 0; k < 3; k++) data->waypoints[i][k] = pos[i][k];
	for (int i = 0; i + 1 < n; i++) {
		real d[3], len = 0;
		for (int k = 0; k < 3; k++) { d[k] = pos[i + 1][k] - pos[i][k]; len += d[k] * d[k]; }
		len = sqrt(len);
		for (int k = 0; k < 3; k++) data->waypoints[i][3 + k] = d[k] / len;
	}
	if (n >= 2)
		for (int k = 3; k < 6; k++) data->waypoints[n - 1][k] = data->waypoints[n - 2][k];
}

obstacle::parameters_struct & obstacle::GetParameterData() { return data->parameters; }

const std::vector< std::vector<real> > & obstacle::GetPath() { return data->waypoints; }

int obstacle::Count() const { return (int)data->type.size(); }

void obstacle::Table(std::vector<real> & type, std::vector<real> & pos, std::vector<real> & rad) const {
	const int n = Count();
	type = data->type;
	pos.resize(3 * n); rad.resize(3 * n);
	for (int i = 0; i < n; i++)
		for (int k = 0; k < 3; k++) {
			pos[3 * i + k] = (i < (int)data->position.size()) ? data->position[i][k] : 0;
			rad[3 * i + k] = (i < (int)data->radius.size()) ? data->radius[i][k] : 0;
		}
}

void obstacle::Upload() const {
	if (data->uploaded) return;
	std::vector<real> type, pos, rad;
	Table(type, pos, rad);
	model::SetDeviceObstacles(Count(), type.data(), pos.data(), rad.data());     // every engine context, present and future
	data->uploaded = true;
}

// map interface: the penalty and its gradient at a position, through a vtolUAV point evaluation on the
// device (H and the costate right-hand side isolate them when the velocity and the costates vanish)
void obstacle::Function(std::vector<real> const& position, real & funcTot) const {
	Upload();
	socp_ctx *ctx = model::Context();
	// vtolUAV.cpp:151-192 with v = 0, p = 0, alphaT = 0, alphaV = 0: H reduces to the map value
	real mp[13] = {10, 0.3, 0, 0, 0, 1, 0, 0, 0, data->parameters.phiObs, data->parameters.psiWP, data->parameters.muObs, data->parameters.sigmaWP};
	real X[12] = {position[0], position[1], position[2], 0, 0, 0, 0, 0, 0, 0, 0, 0}, t = 0, H = 0;
	if (socp_point_batch(ctx, SOCP_VTOL_UAV, 1, mp, nullptr, nullptr, &t, X, nullptr, nullptr, &H, SOCP_HOST) != SOCP_OK) {
		std::cerr << std::endl << "socp_b200: " << socp_last_error(ctx) << std::endl;
		exit(1);
	}
	funcTot = H;
}

void obstacle::Gradient(std::vector<real> const& position, std::vector<real> & gradTot) const {
	Upload();
	socp_ctx *ctx = model::Context();
	// vtolUAV.cpp:58-107: the position costates obey dp/dt = -grad(map) (+ terms that vanish here)
	real mp[13] = {10, 0.3, 0, 0, 0, 1, 0, 0, 0, data->parameters.phiObs, data->parameters.psiWP, data->parameters.muObs, data->parameters.sigmaWP};
	real X[12] = {position[0], position[1], position[2], 0, 0, 0, 0, 0, 0, 0, 0, 0}, t = 0, rhs[12];
	if (socp_point_batch(ctx, SOCP_VTOL_UAV, 1, mp, nullptr, nullptr, &t, X, rhs, nullptr, nullptr, SOCP_HOST) != SOCP_OK) {
		std::cerr << std::endl << "socp_b200: " << socp_last_error(ctx) << std::endl;
		exit(1);
	}
	gradTot.resize(3);
	for (int k = 0; k < 3; k++) gradTot[k] = -rhs[6 + k];
}
