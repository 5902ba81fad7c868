// TEST INFRASTRUCTURE.  Stand-in for <boost/serialization/serialization.hpp>: DBoW2's BowVector / FeatureVector declare a
// serialize() member template that is never instantiated when only transform() is used.
#pragma once
namespace boost { namespace serialization {
class access;
template <class Base, class Derived> inline Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}}
