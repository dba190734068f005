// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's host-only sources (KPR/PZsparse.cu, Trajectory.cu,
// Dynamics.cu) include the CUDA runtime headers through KPR/Headers.h:4-6 but use nothing from them; this
// empty stand-in lets g++ compile those files as C++ (oracle/Makefile.ref).
#pragma once
#include <sys/types.h>  // uint
