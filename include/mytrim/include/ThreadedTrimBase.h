// include/ThreadedTrimBase.h — apps/include/ThreadedTrimBase.h of the reference; see mytrim.h
#ifndef MYTRIM_B200_FWD_APPS_THREADEDTRIMBASE_H
#define MYTRIM_B200_FWD_APPS_THREADEDTRIMBASE_H
#include "../mytrim.h"
#endif
