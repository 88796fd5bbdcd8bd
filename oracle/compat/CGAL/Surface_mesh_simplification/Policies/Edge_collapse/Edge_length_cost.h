// oracle/compat/CGAL/Surface_mesh_simplification/Policies/Edge_collapse/Edge_length_cost.h — TEST INFRASTRUCTURE ONLY: forwards to the stand-in (CGAL is not in this image).
#include <CGAL/mvs_cgal_standin.h>
