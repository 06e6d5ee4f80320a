"""helmholtz_x/parameters_utils.py: coefficient-field builders (O(n) host formulas;
inputs of the hot path, not part of it)."""
import numpy as np

from .fem import DG0Space, Function, functionspace


def gaussian(x, x_ref, sigma, n):
    """parameters_utils.py:8-33."""
    x_ref = np.asarray(x_ref, float)
    if x_ref.ndim == 2 and len(x_ref) == 1:
        x_ref = x_ref[0]
    spatial = sum((x[k] - x_ref[k]) ** 2 for k in range(n))
    amplitude = 1 / (sigma ** n * (2 * np.pi) ** (n / 2))
    return amplitude * np.exp(-1 * spatial / (2 * sigma ** 2))


def _p1_integral(mesh, nodal):
    vol = mesh.volumes().cpu().numpy()
    return float((vol * np.real(nodal)[mesh.cells].mean(axis=1)).sum())


def normalize(func):
    """dolfinx_utils.normalize (dolfinx_utils.py:32-48): divide by int f dx."""
    mesh = func.function_space.mesh
    func.x.array[:] /= _p1_integral(mesh, func.x.array[:mesh.n_nodes])
    return func


def gaussianFunction(mesh, x_r, a_r, degree=1):
    V = functionspace(mesh, ("CG", degree))
    w = Function(V)
    w.interpolate(lambda x: gaussian(x, x_r, a_r, 3))
    return normalize(w)


def halfGaussianFunction(mesh, x_flame, a_flame, degree=1):
    V = functionspace(mesh, ("CG", degree))
    h = gaussianFunction(mesh, x_flame, a_flame, degree=degree)
    xf = np.asarray(x_flame, float).reshape(-1)
    z = V.tabulate_dof_coordinates()[:, 2]
    h.x.array[z < xf[2]] = 0.0
    return normalize(h)


def gamma_function(temperature):
    """parameters_utils.py:62-78."""
    r_gas = 287.1
    if isinstance(temperature, Function):
        gamma = Function(temperature.function_space)
        cp = 973.60091 + 0.1333 * temperature.x.array
        gamma.x.array[:] = cp / (cp - r_gas)
        return gamma
    cp = 973.60091 + 0.1333 * temperature
    return cp / (cp - r_gas)


def sound_speed_variable_gamma(mesh, temperature, degree=1):
    """parameters_utils.py:80-93."""
    c = Function(temperature.function_space, name="soundspeed")
    r_gas = 287.1
    gamma = gamma_function(temperature)
    c.x.array[:] = np.sqrt(gamma.x.array * r_gas * temperature.x.array)
    return c


def sound_speed(temperature):
    c = Function(temperature.function_space, name="soundspeed")
    c.x.array[:] = 20.05 * np.sqrt(temperature.x.array)
    return c


def density_step(x, x_f, sigma, rho_d, rho_u):
    return rho_u + (rho_d - rho_u) / 2 * (1 + np.tanh((x - x_f) / (sigma)))


def rho_step(mesh, x_f, a_f, rho_d, rho_u, degree=1):
    V = functionspace(mesh, ("CG", degree))
    rho = Function(V)
    zf = np.asarray(x_f, float).reshape(-1)[2]
    rho.interpolate(lambda x: density_step(x[2], zf, a_f, rho_d, rho_u))
    return rho


def rho_ideal(temperature, p_0, r_gas):
    density = Function(temperature.function_space)
    density.x.array[:] = p_0 / (r_gas * temperature.x.array)
    return density


def _step(mesh, x_f, up, down, degree, name):
    V = functionspace(mesh, ("CG", degree))
    f = Function(V, name=name)
    zf = np.asarray(x_f, float).reshape(-1)[2]
    z = V.tabulate_dof_coordinates()[:, 2]
    f.x.array[:] = np.where(z < zf, up, down)
    return f


def c_step(mesh, x_f, c_u, c_d):
    return _step(mesh, x_f, c_u, c_d, 1, "soundspeed")


def temperature_step(mesh, x_f, T_u, T_d, degree=1):
    return _step(mesh, x_f, T_u, T_d, degree, "temperature")


def c_uniform(mesh, sos, degree=1):
    f = Function(functionspace(mesh, ("CG", degree)), name="soundspeed")
    f.x.array[:] = sos
    return f


def temperature_uniform(mesh, temp):
    f = Function(functionspace(mesh, ("CG", 1)), name="temperature")
    f.x.array[:] = temp
    return f


def temperature(mesh, soundSpeed):
    T = Function(functionspace(mesh, ("CG", 1)), name="temperature")
    ss = soundSpeed.x.array if isinstance(soundSpeed, Function) else soundSpeed
    T.x.array[:] = np.square(ss) / (287.1 * 1.4)
    return T


def Q_volumetric(mesh, subdomains, Q_total, flame_tag=0, degree=0):
    q = Function(DG0Space(mesh))
    vol = mesh.volumes().cpu().numpy()
    sel = mesh.cell_tags == flame_tag
    q.x.array[sel] = Q_total / vol[sel].sum()
    return q


def Q_multiple(mesh, subdomains, N_sector, degree=0):
    """parameters_utils.py:228-247: DG0, 1/V_f on the cells tagged f.  Computed on the device (masked
    sums, fixed reduction order); the host array appears only if the caller reads it."""
    import torch
    vol = mesh.volumes()
    tags = mesh.cell_tagsd
    q = torch.zeros_like(vol)
    for flame in range(N_sector):
        sel = tags == flame
        q = torch.where(sel, 1.0 / vol[sel].sum(), q)
    return Function.from_device(DG0Space(mesh), q)
