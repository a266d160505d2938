#pragma once
#include <limits>
#define CUDART_INF (std::numeric_limits<double>::infinity())
#define CUDART_INF_F (std::numeric_limits<float>::infinity())
#define CUDART_NAN (std::numeric_limits<double>::quiet_NaN())
