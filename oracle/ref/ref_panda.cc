// TEST INFRASTRUCTURE ONLY: instantiates the harness for the reference's generated Panda program
// (src/impl/vamp/robots/panda.hh), included from /root/reference, never copied.
#include <vamp/robots/panda.hh>
#include "ref_robot.hh"

extern const refh::RobotVTable ref_vt_panda = refh::Harness<vamp::robots::Panda>::vtable;
