// Stand-alone (unfused) kernels behind the reference's layers.py module API.
#pragma once
#include "common.cuh"
