// optimet3d_b200 -- command-line driver, same invocation as the reference (`Optimet3D input.xml`,
// srcAna/main.cpp:23-46): runs the response scan (the reference's .dat files) or the field map on the B200 path.
#include "ob_host.hpp"
#include <iostream>

int main(int argc, char *argv[]) {
  if(argc <= 1) {
    std::cerr << "Please give an input file (without the .xml extension), as in: " << argv[0] << " case" << std::endl;
    return 1;
  }
  try {
    std::string caseFile = argv[1];
    if(caseFile.size() > 4 && caseFile.substr(caseFile.size() - 4) == ".xml")
      caseFile = caseFile.substr(0, caseFile.size() - 4);
    optimet_b200::Run run = optimet_b200::simulation_input(caseFile + ".xml"); // Simulation.cpp:34-46
    optimet_b200::solver::B200Matrix solver(run, argc > 2 ? std::atoi(argv[2]) : 0);
    if(run.outputType == 0) { // Simulation.cpp:48-53
      optimet_b200::FieldMap fm = optimet_b200::field_simulation(run, solver, caseFile);
      std::cout << "Field map " << fm.nx << " x " << fm.ny << " x " << fm.nz << " written to " << caseFile
                << "_FF.field" << (run.excitation->SH_cond ? " and _SH.field" : "") << std::endl;
      return 0;
    }
    if(run.outputType != 11) {
      std::cerr << "Only <output type=\"response\"> wavelength scans and <output type=\"field\"> maps run on the B200 path"
                << std::endl;
      return 2;
    }
    std::vector<optimet_b200::ScanLine> lines = optimet_b200::scan_wavelengths(run, solver, caseFile);
    for(auto const &l : lines)
      std::cout << "Lambda = " << l.lambda << "  sca_FF = " << l.scattering_FF << "  abs_FF = " << l.absorption_FF
                << "  sca_SH = " << l.scattering_SH << "  abs_SH = " << l.absorption_SH << "  GMRES it = " << l.iters_FF
                << "/" << l.iters_SH << std::endl;
  } catch(std::exception &e) {
    std::cerr << "Error: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
