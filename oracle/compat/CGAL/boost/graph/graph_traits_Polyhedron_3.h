// oracle/compat/CGAL/boost/graph/graph_traits_Polyhedron_3.h — TEST INFRASTRUCTURE ONLY: forwards to the stand-in (CGAL is not in this image).
#include <CGAL/mvs_cgal_standin.h>
