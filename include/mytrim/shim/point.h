// shim/point.h — see mytrim.h
#ifndef MYTRIM_B200_FWD_SHIM_POINT_H
#define MYTRIM_B200_FWD_SHIM_POINT_H
#include "../mytrim.h"
#endif
