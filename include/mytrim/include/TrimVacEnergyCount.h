// include/TrimVacEnergyCount.h — apps/include/TrimVacEnergyCount.h of the reference; see mytrim.h
#ifndef MYTRIM_B200_FWD_APPS_TRIMVACENERGYCOUNT_H
#define MYTRIM_B200_FWD_APPS_TRIMVACENERGYCOUNT_H
#include "../mytrim.h"
#endif
