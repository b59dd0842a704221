// Forwarding header: the samurai API for the hot path lives in b200_api.hpp.
#pragma once
#include "../b200_api.hpp"
