// include/TrimRange.h — apps/include/TrimRange.h of the reference; see mytrim.h
#ifndef MYTRIM_B200_FWD_APPS_TRIMRANGE_H
#define MYTRIM_B200_FWD_APPS_TRIMRANGE_H
#include "../mytrim.h"
#endif
