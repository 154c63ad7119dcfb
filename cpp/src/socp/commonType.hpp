// cpp/src/socp/commonType.hpp -- mirror of src/socp/commonType.hpp:8-18 (fp64 only; the float
// build of the reference is a commented-out switch and is out of scope).
#ifndef SOCP_B200_COMMONTYPE_HPP
#define SOCP_B200_COMMONTYPE_HPP
typedef double real;
#endif
