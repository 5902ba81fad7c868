// TEST INFRASTRUCTURE.  Stand-in for <boost/serialization/access.hpp> (see serialization.hpp next to it).
#pragma once
#include "serialization.hpp"
