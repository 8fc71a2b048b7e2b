// TEST INFRASTRUCTURE ONLY: instantiates the harness for the reference's generated UR5 program
// (src/impl/vamp/robots/ur5.hh), included from /root/reference, never copied.
#include <vamp/robots/ur5.hh>
#include "ref_robot.hh"

extern const refh::RobotVTable ref_vt_ur5 = refh::Harness<vamp::robots::UR5>::vtable;
