Three-group test library for the mocc_b200 golden fixtures (synthetic values)
 3 3
 1.0E+07 1.0E+03 0.625
!
! Synthetic 3-group macroscopic cross sections, written for these tests only.
! Columns: absorption, nu-fission, fission, chi; then the scattering matrix
! (row = destination group, column = source group), with some up-scatter from
! group 3 to group 2.
!
XSMACRO fuelA 0
  9.100E-03 7.300E-03 2.900E-03 7.50E-01
  2.850E-02 2.100E-02 8.600E-03 2.50E-01
  1.130E-01 1.720E-01 7.050E-02 0.00E+00
  2.1500E-01 0.0000E+00 0.0000E+00
  1.8300E-02 3.5200E-01 2.1000E-03
  3.2000E-05 9.4000E-03 4.1100E-01

XSMACRO fuelB 0
  1.020E-02 8.900E-03 3.300E-03 7.50E-01
  3.470E-02 2.900E-02 1.150E-02 2.50E-01
  1.580E-01 2.450E-01 9.800E-02 0.00E+00
  2.1200E-01 0.0000E+00 0.0000E+00
  1.7100E-02 3.4400E-01 2.6000E-03
  2.8000E-05 8.1000E-03 4.0200E-01

XSMACRO water 0
  4.500E-04 0.000E+00 0.000E+00 0.00E+00
  2.300E-03 0.000E+00 0.000E+00 0.00E+00
  3.100E-02 0.000E+00 0.000E+00 0.00E+00
  1.9800E-01 0.0000E+00 0.0000E+00
  6.3000E-02 5.7100E-01 4.8000E-03
  1.1000E-03 1.2400E-01 1.9200E+00
