"""CPU oracle — test infrastructure only (see oracle/qg_oracle.hpp).  Product code never imports this."""
