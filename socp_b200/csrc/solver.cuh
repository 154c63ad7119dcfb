// placeholder, replaced by the batched solver
#pragma once
namespace socp {
struct SolverWorkspace {
    void release() {}
    double bytes() const { return 0; }
};
}
