// forwarding header: the reference API of include/vec3.h lives in particle_simulation_b200.hpp
#pragma once
#include "../particle_simulation_b200.hpp"
