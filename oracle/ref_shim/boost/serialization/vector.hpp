// TEST INFRASTRUCTURE.  Stand-in for <boost/serialization/vector.hpp> (see serialization.hpp next to it).
#pragma once
#include "serialization.hpp"
