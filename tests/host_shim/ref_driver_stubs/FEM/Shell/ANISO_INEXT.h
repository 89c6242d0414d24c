#pragma once
// TEST INFRASTRUCTURE: shadows the reference header of the same name (strain limiting / fibers need Eigen::JacobiSVD and FullPivLU,
// which the Eigen stand-in does not have); tests/host_shim/ref_driver.cpp provides aborting stand-ins for its functions.
