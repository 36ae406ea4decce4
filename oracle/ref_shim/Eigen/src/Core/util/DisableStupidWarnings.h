// oracle/ref_shim — TEST INFRASTRUCTURE: intentionally empty (PoseAdapterBase.hpp:13 includes it).
