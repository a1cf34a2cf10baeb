// stand-in for glog: LOG(severity) << ... goes to std::cerr (test infrastructure, see ../README.md)
#ifndef SIPP_SHIM_GLOG_H_
#define SIPP_SHIM_GLOG_H_
#include <iostream>
#define LOG(severity) std::cerr
#endif
