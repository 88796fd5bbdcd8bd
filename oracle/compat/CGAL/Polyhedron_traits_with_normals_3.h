// oracle/compat/CGAL/Polyhedron_traits_with_normals_3.h — TEST INFRASTRUCTURE ONLY: forwards to the stand-in (CGAL is not in this image).
#include <CGAL/mvs_cgal_standin.h>
