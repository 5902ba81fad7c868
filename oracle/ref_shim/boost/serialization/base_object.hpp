// TEST INFRASTRUCTURE.  Stand-in for <boost/serialization/base_object.hpp> (see serialization.hpp next to it).
#pragma once
#include "serialization.hpp"
