// ORACLE SUPPORT (test infrastructure): what cmake's configure_file makes of Thirdparty/g2o/config.h.in with OpenMP off
// (CMakeLists.txt: G2O_USE_OPENMP OFF) and shared libraries on.  Found as "../../config.h" through -Iref_shim/g2o_cfg/x/y.
#ifndef G2O_CONFIG_H
#define G2O_CONFIG_H
#define G2O_SHARED_LIBS 1
#endif
