#pragma once
/* TEST INFRASTRUCTURE: see tensorflow/tf_stub.h */
#include "tensorflow/tf_stub.h"
