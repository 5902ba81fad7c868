// TEST INFRASTRUCTURE.  Stand-in for <boost/serialization/map.hpp> (nothing needed, see serialization.hpp next to it).
#pragma once
