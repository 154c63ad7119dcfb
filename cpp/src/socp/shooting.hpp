// cpp/src/socp/shooting.hpp -- mirror of the reference's `shooting` class
// (src/socp/shooting.hpp:18-383): same public API, solved on the B200 through the C ABI
// (socp_solve_batch with a batch of one; continuation loops as shooting.cpp:598-778).
#include <iostream>
#include <unordered_map>

#include "commonType.hpp"
#include "model.hpp"

#ifndef _SHOOTING_H_
#define _SHOOTING_H_

class shooting
{
public:
	shooting(model & model, int numMulti = int(1), int numThread = int(1));
	~shooting();

	void Resize(int numMulti, int numThread) const;
	void SetMode(int const& mode_tf, std::vector<int> const& mode_Xf) const;
	void SetMode(std::vector<int> const& mode_t, std::vector< std::vector<int> > const& mode_X) const;
	void InitShooting(real const& ti, model::mstate const& Xi, real const& tf, model::mstate const& Xf) const;
	void InitShooting(std::vector<real> const& vt, std::vector<model::mstate> const& vX) const;
	void SetDesiredState(real const& ti, model::mstate const& Xi, real const& tf, model::mstate const& Xf) const;
	void SetDesiredState(std::vector<real> const& vt, std::vector<model::mstate> const& vX) const;

	int SolveOCP(real const& continuationStep) const;
	/// the reference runs the solve on a helper thread and abandons it after timeoutMS
	/// (shooting.cpp:329-348); here the batched solve is bounded by maxfev instead and the
	/// timeout is ignored
	int SolveOCP(real const& continuationStep, double const& timeoutMS) const;
	int SolveOCP(real const& continuationStep, real & Rdata, real const& Rgoal) const;
	int SolveOCP(real const& continuationStep, std::string const& Rdata, real const& Rgoal) const {
		if (myModel.parameters.find(Rdata) != myModel.parameters.end()) {
			return SolveOCP(continuationStep, myModel.parameters[Rdata], Rgoal);
		}
		else {
			std::cout << std::endl << std::endl << "Data " << Rdata << " does not exist!" << std::endl << std::endl;
		}
		return 0;
	};

	model::mstate Move(real const& ti, model::mstate const& Xi, real const& tf, int isJac = 0) const;
	void Move(real const& ti, model::mstate const& Xi, real const& tf, model::mstate & Xf, int isJac = 0) const;
	model::mstate Move(real const& tf, int isJac = 0) const;
	void Move(real const& tf, model::mstate & Xf, int isJac = 0) const;

	void SetPrecision(real const& xtol) const;
	void SetContinuationMinStep(real const& step) const;
	real GetParameters(int const& k) const;
	real *GetParameters() const;
	void GetParameters(std::vector<real> & paramVector) const;
	void GetSolution(std::vector<real> & vt, std::vector<model::mstate> & vX) const;
	std::vector<int> GetCallNumber() const;
	void Trace() const;
	model & GetModel() const;

	// ---- B200 engine extras ---------------------------------------------------------------------
	/// norm of the residual at the last solution (the reference only exposes `info`)
	real GetResidualNorm() const;

private:
	model & myModel;
	struct data_struct;
	data_struct *data;

	int SolveShooting() const;
	int SolveShootingContinuation(real const& continuationStep) const;
	int SolveShootingContinuation(real const& continuationStep, real & Rdata, real const& Rgoal) const;
	int SolveShootingFunction(int const & numParam, std::vector<real> & param) const;
	void UpdateSolution() const;
	void ComputeTimeLine(std::vector<real> const& param, std::vector<real> & timeLine) const;
};

#endif //_SHOOTING_H_
