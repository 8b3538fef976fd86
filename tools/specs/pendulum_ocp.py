# OCP description of the second model (pendulum on a cart, nx = 4, nu = 1) in acados_template terms, for tools/gen_spec.py
# (another nx, nu through the same generator and the same kernel sources).  The dynamics are the reference's
# own model file, acados/examples/acados_python/pendulum_on_cart/common/pendulum_model.py; horizon, cost, bounds and
# initial state are those of acados/examples/acados_python/tests/test_ocp_setting.py:150-205 (N = 20, Tf = 1,
# Q = 2 diag(1e3, 1e3, 1e-2, 1e-2), R = 2 diag(1e-2), W_e = Q, |F| <= 80, x0 = (0, pi, 0, 0)), with the integrator and
# solver choices of the Crazyflie OCP (one ERK4 step per interval, SQP_RTI, Gauss-Newton, partial-condensing HPIPM).
from acados_template import AcadosOcp
from pendulum_model import export_pendulum_ode_model
import numpy as np

ocp = AcadosOcp()
model = export_pendulum_ode_model()
ocp.model = model
nx, nu = 4, 1
ny, ny_e = nx + nu, nx
N = 20
ocp.dims.N = N
Q = 2 * np.diag([1e3, 1e3, 1e-2, 1e-2])
R = 2 * np.diag([1e-2])
W = np.zeros((ny, ny))
W[:nx, :nx] = Q
W[nx:, nx:] = R
ocp.cost.cost_type = 'LINEAR_LS'
ocp.cost.cost_type_e = 'LINEAR_LS'
ocp.cost.W = W
ocp.cost.W_e = Q
Vx = np.zeros((ny, nx))
Vx[:nx, :nx] = np.eye(nx)
Vu = np.zeros((ny, nu))
Vu[nx, 0] = 1.0
ocp.cost.Vx = Vx
ocp.cost.Vu = Vu
ocp.cost.Vx_e = np.eye(nx)
ocp.cost.yref = np.zeros((ny,))
ocp.cost.yref_e = np.zeros((ny_e,))
Fmax = 80
ocp.constraints.lbu = np.array([-Fmax])
ocp.constraints.ubu = np.array([+Fmax])
ocp.constraints.idxbu = np.array([0])
ocp.constraints.x0 = np.array([0.0, np.pi, 0.0, 0.0])
ocp.solver_options.qp_solver = 'PARTIAL_CONDENSING_HPIPM'
ocp.solver_options.hessian_approx = 'GAUSS_NEWTON'
ocp.solver_options.integrator_type = 'ERK'
ocp.solver_options.nlp_solver_type = 'SQP_RTI'
ocp.solver_options.tf = 1.0
