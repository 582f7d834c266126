// Stand-in for a Boost.uBLAS header -- TEST INFRASTRUCTURE ONLY.  /root/reference/ACFcalculator.cpp includes it (:5-7)
// and uses nothing from it.
#pragma once
