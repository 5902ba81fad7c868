// TEST INFRASTRUCTURE.  Stand-in for <boost/serialization/assume_abstract.hpp> (see serialization.hpp next to it).
#pragma once
#include "serialization.hpp"
