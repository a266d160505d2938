"""Import-path shim for the reference's blind-detection package (PolarEncoder/PolarBD): the same module layout --
PolarBD.PolarBD.DMetricCalculator / PolarBD.PolarBD.CASCLWithRNTI over PolarBD._cpp.libPolarBD -- on the B200 build."""
