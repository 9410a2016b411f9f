#pragma once
#include "discrete_grid.hpp"
