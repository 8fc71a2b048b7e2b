// TEST INFRASTRUCTURE ONLY: instantiates the harness for the reference's generated Fetch program
// (src/impl/vamp/robots/fetch.hh), included from /root/reference, never copied.
#include <vamp/robots/fetch.hh>
#include "ref_robot.hh"

extern const refh::RobotVTable ref_vt_fetch = refh::Harness<vamp::robots::Fetch>::vtable;
