// TEST INFRASTRUCTURE ONLY: see ../Graphics.hpp
#pragma once
#include "../Graphics.hpp"
