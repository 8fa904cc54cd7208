// ORACLE — TEST INFRASTRUCTURE ONLY.  util/settings.cpp includes <boost/bind.hpp> without using it.
#pragma once
