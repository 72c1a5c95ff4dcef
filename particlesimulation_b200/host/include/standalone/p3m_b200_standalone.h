// marker: this directory is on the include path, i.e. the caller builds WITHOUT the reference checkout.
// particle_simulation_b200.hpp then defines Vec3, Particle, the pmConfig enums, StateRecorder, FFTAdapter,
// AbstractStepper and the external-field functions itself instead of including the reference's headers.
#pragma once
#define P3M_B200_STANDALONE 1
