// Forwarding header of the oracle-only mini cv shim (see ../opencv.hpp).
#pragma once
#include "../opencv.hpp"
