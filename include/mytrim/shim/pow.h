// shim/pow.h — see mytrim.h
#ifndef MYTRIM_B200_FWD_SHIM_POW_H
#define MYTRIM_B200_FWD_SHIM_POW_H
#include "../mytrim.h"
#endif
