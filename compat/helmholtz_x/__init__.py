"""Import-name shim: lets a helmholtz-x driver written for the reference
(`from helmholtz_x.eigensolvers import fixed_point_iteration` ...) run on helmholtz_x_b200
when this directory is put on PYTHONPATH.  See INTEGRATION.md."""
