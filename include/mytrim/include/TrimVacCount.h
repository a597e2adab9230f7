// include/TrimVacCount.h — apps/include/TrimVacCount.h of the reference; see mytrim.h
#ifndef MYTRIM_B200_FWD_APPS_TRIMVACCOUNT_H
#define MYTRIM_B200_FWD_APPS_TRIMVACCOUNT_H
#include "../mytrim.h"
#endif
