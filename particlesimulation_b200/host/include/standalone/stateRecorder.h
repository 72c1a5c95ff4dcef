// forwarding header: the reference API of include/stateRecorder.h lives in particle_simulation_b200.hpp
#pragma once
#include "../particle_simulation_b200.hpp"
