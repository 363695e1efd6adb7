// TEST INFRASTRUCTURE ONLY -- placeholder for <nlopt.hpp> (gple/stdafx.h:52) when the translation
// units that do not optimise (kernel, complex_kernel, pes, evolve, predict, mc) are compiled into
// oracle/_ref/.  opt.cpp is compiled against oracle/refstub/nlopt_full.hpp instead (see Makefile.ref).
#pragma once
#ifdef REFSTUB_WITH_NLOPT
#include "nlopt_full.hpp"
#endif
