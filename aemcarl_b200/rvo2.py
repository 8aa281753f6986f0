"""`rvo2.PyRVOSimulator`-shaped adapter over the sm_100a ORCA kernel (aem_orca).

The reference talks to ORCA through Python-RVO2 (crowd_sim/envs/policy/orca.py:95-129, crowd_sim.py:456-480):
    sim = rvo2.PyRVOSimulator(timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed)
    sim.addAgent(pos, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity)
    sim.setAgentPosition / setAgentVelocity / setAgentPrefVelocity(i, v);  sim.doStep();  sim.getAgentVelocity(i)
This class keeps that surface.  doStep() solves every agent's ORCA program against all the others with ONE aem_orca launch
(one warp per agent), then advances velocities and positions like RVOSimulator::doStep.  There is no CPU path.

Scope = the parameter set the reference uses: neighborDist 10, maxNeighbors 10, timeHorizon 5, no obstacles, at most 32
agents, preferred velocities of norm <= 1 (orca.py:113-115 normalises them; the kernel re-derives the preferred velocity
as `goal - position`, normalised when longer than 1).  Other values raise NotImplementedError instead of silently
computing something else.
"""
import numpy as np
import torch


class PyRVOSimulator(object):
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity=(0, 0)):
        self._check(neighborDist, maxNeighbors, timeHorizon)
        self.time_step = float(timeStep)
        self.default_radius, self.default_max_speed = float(radius), float(maxSpeed)
        self.pos, self.vel, self.pref, self.radius, self.max_speed = [], [], [], [], []
        self._engine = None

    @staticmethod
    def _check(neighborDist, maxNeighbors, timeHorizon):
        if float(neighborDist) != 10.0 or int(maxNeighbors) != 10 or float(timeHorizon) != 5.0:
            raise NotImplementedError("aem_orca implements the reference's ORCA parameters (neighborDist 10, maxNeighbors 10, "
                                      "timeHorizon 5; orca.py:61-66)")

    # ---- agents ---------------------------------------------------------------------------------
    def addAgent(self, pos, neighborDist=10, maxNeighbors=10, timeHorizon=5, timeHorizonObst=5, radius=None, maxSpeed=None,
                 velocity=(0, 0)):
        self._check(neighborDist, maxNeighbors, timeHorizon)
        if len(self.pos) >= 32:
            raise NotImplementedError("at most 32 agents (AEM_MAX_HUMANS)")
        self.pos.append([float(pos[0]), float(pos[1])])
        self.vel.append([float(velocity[0]), float(velocity[1])])
        self.pref.append([0.0, 0.0])
        self.radius.append(self.default_radius if radius is None else float(radius))
        self.max_speed.append(self.default_max_speed if maxSpeed is None else float(maxSpeed))
        return len(self.pos) - 1

    def getNumAgents(self):
        return len(self.pos)

    def setAgentPosition(self, i, p):
        self.pos[i] = [float(p[0]), float(p[1])]

    def setAgentVelocity(self, i, v):
        self.vel[i] = [float(v[0]), float(v[1])]

    def setAgentPrefVelocity(self, i, v):
        self.pref[i] = [float(v[0]), float(v[1])]

    def getAgentPosition(self, i):
        return tuple(self.pos[i])

    def getAgentVelocity(self, i):
        return tuple(self.vel[i])

    def getAgentPrefVelocity(self, i):
        return tuple(self.pref[i])

    # ---- one simulation step ---------------------------------------------------------------------
    def doStep(self):
        from aemcarl_b200.engine import get_engine, make_env_params
        n = len(self.pos)
        if n == 0:
            return
        if self._engine is None:
            self._engine = get_engine(torch.cuda.current_device(), role="rvo2")
        eng = self._engine
        pref = np.asarray(self.pref, dtype=np.float64)
        if (np.hypot(pref[:, 0], pref[:, 1]) > 1.0 + 1e-12).any():
            raise NotImplementedError("preferred velocities longer than 1 are outside the reference's call pattern "
                                      "(orca.py:113-115 normalises them)")
        humans = np.zeros((1, n, 8), dtype=np.float64)
        humans[0, :, 0:2] = self.pos
        humans[0, :, 2:4] = self.vel
        humans[0, :, 4] = np.asarray(self.radius) - 0.01       # the kernel adds the reference's 0.01 margin itself (orca.py:100)
        humans[0, :, 5:7] = humans[0, :, 0:2] + pref           # prefVel = goal - position
        humans[0, :, 7] = self.max_speed
        robot = np.zeros((1, 9), dtype=np.float64)
        eng.set_env_params(make_env_params(time_step=self.time_step, robot_visible=False, orca_safety_space=0.0))
        nv = eng.orca(torch.from_numpy(humans).to(eng.device), torch.from_numpy(robot).to(eng.device))[0].cpu().numpy()
        for i in range(n):                                      # RVOSimulator::doStep: v <- v', p += v * dt
            self.vel[i] = [float(nv[i, 0]), float(nv[i, 1])]
            self.pos[i] = [self.pos[i][0] + self.vel[i][0] * self.time_step, self.pos[i][1] + self.vel[i][1] * self.time_step]
