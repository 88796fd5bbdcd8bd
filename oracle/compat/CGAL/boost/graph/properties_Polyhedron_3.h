// oracle/compat/CGAL/boost/graph/properties_Polyhedron_3.h — TEST INFRASTRUCTURE ONLY: forwards to the stand-in (CGAL is not in this image).
#include <CGAL/mvs_cgal_standin.h>
