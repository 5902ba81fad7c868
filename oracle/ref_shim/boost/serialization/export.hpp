// TEST INFRASTRUCTURE.  Stand-in for <boost/serialization/export.hpp> (see serialization.hpp next to it).
#pragma once
#include "serialization.hpp"
