// zk_host.cpp -- API-mirror half of the C ABI (placeholder, filled in below)
#include "../../include/zeekstd_b200.h"
