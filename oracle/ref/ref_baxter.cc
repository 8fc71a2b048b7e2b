// TEST INFRASTRUCTURE ONLY: instantiates the harness for the reference's generated Baxter program
// (src/impl/vamp/robots/baxter.hh), included from /root/reference, never copied.
#include <vamp/robots/baxter.hh>
#include "ref_robot.hh"

extern const refh::RobotVTable ref_vt_baxter = refh::Harness<vamp::robots::Baxter>::vtable;
