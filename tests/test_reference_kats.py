"""The reference's own known-answer tests for this path, run against the ORACLE (pins the oracle beside the golden
rollouts; SURVEY.md section 8c lists them).  tests/test_gpu_reference_kats.py runs the same scenarios on the GPU."""
import ctypes as C

import numpy as np
import pytest

import hwy_oracle as ho
import kat_scenarios as K
import net_oracle as no


def test_rotated_rectangles_intersect():
    """tests/test_utils.py:19-27"""
    for (c1, l1, w1, a1), (c2, l2, w2, a2), want in K.RECTANGLES:
        got = no.lib().net_rotated_rectangles_intersect(c1[0], c1[1], l1, w1, a1, c2[0], c2[1], l2, w2, a2)
        assert bool(got) == want


@pytest.mark.parametrize("case", K.NEIGHBOURS, ids=[c[0] for c in K.NEIGHBOURS])
def test_neighbour_vehicles(case):
    """tests/road/test_neighbour_vehicles.py (lines in kat_scenarios.NEIGHBOURS)"""
    net, x, y, h, lanes, query = K.build_neighbour_case(case)
    graph = no.graph_from_arrays(net.export_arrays())
    cfg = no.NetCfg()
    cfg.n_vehicles = len(x)
    cfg.connected_lanes = int(case[3])
    ob = no.NetOracleBatch(graph, cfg, 1)
    ob.a["x"][0], ob.a["y"][0], ob.a["heading"][0], ob.a["lane"][0] = x, y, h, lanes
    ob.a["speed"][0] = 10.0
    st = ob._state(0)
    f, r = C.c_int32(), C.c_int32()
    no.lib().net_neighbours(C.byref(graph), C.byref(cfg), C.byref(st), 0, int(query), C.byref(f), C.byref(r))
    K.check_neighbour_result(case, f.value, r.value)


# ---------------------------------------------------------------- vehicle dynamics / control on the highway oracle
def _highway(lanes, action, V=1, **over):
    from highwayenv_b200.config import default_config

    cfg = default_config("highway-v0")
    cfg.update({"lanes_count": lanes, "vehicles_count": V - 1, "action": action, "duration": 1000})
    cfg.update(over)
    cfg["_others_check_collisions"] = 1
    oc = ho.cfg_from_dict(cfg)
    ob = ho.OracleBatch(oc, 1, seeds=[0])
    ob.reset()
    return ob


def _place(ob, v, x, y, speed, heading=0.0, lane=0):
    a = ob.a
    a["x"][0, v], a["y"][0, v], a["speed"][0, v], a["heading"][0, v] = x, y, speed, heading
    a["lane"][0, v] = a["target_lane"][0, v] = lane
    a["target_speed"][0, v] = speed
    a["crashed"][0, v] = a["has_impact"][0, v] = 0


CONT = {"type": "ContinuousAction"}


def _cont(accel, steer):  # inverse of ContinuousAction's lmap (action.py:136-152)
    return np.array([[accel / 5.0, steer / (np.pi / 4)]], dtype=np.float32)


def test_dynamics_step():
    """vehicle/test_dynamics.py:12-19 (and test_control.py:11-18): 2 s at 20 m/s -> x = 40"""
    ob = _highway(1, CONT)
    _place(ob, 0, 0.0, 0.0, 20.0)
    for _ in range(2):  # one policy step = FPS substeps of 1/FPS
        ob.step(_cont(0, 0))
    assert ob.a["x"][0, 0] == pytest.approx(40) and ob.a["y"][0, 0] == pytest.approx(0)
    assert ob.a["speed"][0, 0] == pytest.approx(20) and ob.a["heading"][0, 0] == pytest.approx(0)


def test_dynamics_act():
    """vehicle/test_dynamics.py:22-33: 1 m/s2 for 1 s -> 21 m/s; then steering 0.5: speed kept, y > 0"""
    ob = _highway(1, CONT)
    _place(ob, 0, 0.0, 0.0, 20.0)
    ob.step(_cont(1, 0))
    assert ob.a["speed"][0, 0] == pytest.approx(21)
    ob.step(_cont(0, 0.5))
    assert ob.a["speed"][0, 0] == pytest.approx(21) and ob.a["y"][0, 0] > 0


def test_dynamics_brake():
    """vehicle/test_dynamics.py:36-41: a = clip(-v, +-6) for 10 s -> standstill (the command is refreshed per
    policy step here: 10 steps of 1 s, same fixed point)"""
    ob = _highway(1, CONT, policy_frequency=15)
    _place(ob, 0, 0.0, 0.0, 20.0)
    for _ in range(10 * K.FPS):
        v = ob.a["speed"][0, 0]
        ob.step(_cont(min(max(-1 * v, -5), 5), 0))  # the action space caps |a| at 5 (action.py:82)
    assert ob.a["speed"][0, 0] == pytest.approx(0, abs=0.01)


def test_dynamics_collision():
    """vehicle/test_dynamics.py:51-60: two vehicles 4 m apart are both crashed after handle_collisions (the
    lane_distance_to = +-10 assertions of :44-49 are checked through the Kinematics row in the GPU twin)"""
    ob = _highway(1, CONT, V=2)
    _place(ob, 0, 0.0, 0.0, 10.0)
    _place(ob, 1, 4.0, 0.0, 20.0)
    ob.step(_cont(0, 0))
    assert ob.a["crashed"][0, 0] and ob.a["crashed"][0, 1]


META = {"type": "DiscreteMetaAction", "target_speeds": [20, 25, 30]}


def test_control_lane_change():
    """vehicle/test_control.py:21-38: LANE_RIGHT, 3 s later the vehicle sits on lane 1 at 20 m/s"""
    ob = _highway(2, META)
    _place(ob, 0, 0.0, 0.0, 20.0)
    ob.a["kind"][0, 0] = 1
    ob.a["speed_index"][0] = 0
    ob.step(np.array([2]))
    for _ in range(2):
        ob.step(np.array([1]))
    assert ob.a["speed"][0, 0] == pytest.approx(20)
    assert ob.a["y"][0, 0] == pytest.approx(4.0, abs=1.0) and ob.a["lane"][0, 0] == 1


def test_control_speed():
    """vehicle/test_control.py:41-54: FASTER, 3 tau later the speed reached 20 + DELTA_SPEED within 0.5"""
    ob = _highway(1, META)
    _place(ob, 0, 0.0, 0.0, 20.0)
    ob.a["kind"][0, 0] = 1
    ob.a["speed_index"][0] = 0
    ob.step(np.array([3]))
    ob.step(np.array([1]))  # 2 s >= 3 * TAU_ACC = 1.8 s
    assert ob.a["speed"][0, 0] == pytest.approx(25, abs=0.5)
    assert ob.a["y"][0, 0] == pytest.approx(0) and ob.a["lane"][0, 0] == 0


def test_road_network_follow_road():
    """road/test_road.py:23-40 test_network: a ControlledVehicle(target_speed=2) dropped at [5, 0] on the diamond
    network is on lane (0, 1, 0); driving 20 s without a route its target lane changes at least 3 times
    (follow_road -> next_lane's closest-next-road rule)."""
    net = K.diamond_network()
    graph = no.graph_from_arrays(net.export_arrays())
    assert no.lib().net_closest_lane(C.byref(graph), 5.0, 0.0, 0.0) == net.index[("0", "1", 0)]
    cfg = no.NetCfg()
    cfg.n_vehicles, cfg.simulation_frequency, cfg.policy_frequency = 1, 15, 1
    cfg.acc_max, cfg.comfort_acc_max, cfg.comfort_acc_min = 6.0, 3.0, -5.0
    cfg.distance_wanted, cfg.time_wanted, cfg.lane_change_delay = 10.0, 1.5, 1.0
    ob = no.NetOracleBatch(graph, cfg, 1)
    a = ob.a
    a["x"][0, 0], a["y"][0, 0], a["speed"][0, 0], a["target_speed"][0, 0] = 5.0, 0.0, 0.0, 2.0
    a["kind"][0, 0], a["delta"][0, 0] = no.KIND_MDP, 4.0
    st = ob._state(0)
    lane, changes = int(a["target_lane"][0, 0]), 0
    for _ in range(int(20 * 15)):
        no.lib().net_substeps(C.byref(graph), C.byref(cfg), C.byref(st), 1)
        if int(a["target_lane"][0, 0]) != lane:
            lane, changes = int(a["target_lane"][0, 0]), changes + 1
    assert changes >= 3


def _single_lane_oracle(n_slots):
    net = K.single_lane_road()
    graph = no.graph_from_arrays(net.export_arrays())
    cfg = no.NetCfg()
    cfg.n_vehicles, cfg.simulation_frequency, cfg.policy_frequency = n_slots, 15, 1
    cfg.acc_max, cfg.comfort_acc_max, cfg.comfort_acc_min = 6.0, 3.0, -5.0  # IDMVehicle class constants
    cfg.distance_wanted, cfg.time_wanted, cfg.lane_change_delay = 10.0, 1.5, 1.0
    cfg.lane_change_min_acc_gain, cfg.lane_change_max_braking_imposed = 0.2, 2.0
    return graph, cfg, no.NetOracleBatch(graph, cfg, 1)


def test_behavior_stop_before_obstacle():
    """vehicle/test_behavior.py:13-27 (IDMVehicle): a vehicle at 20 m/s stops DISTANCE_WANTED before an Obstacle
    80 m ahead, without crashing"""
    graph, cfg, ob = _single_lane_oracle(2)
    a = ob.a
    a["speed"][0, 0], a["target_speed"][0, 0], a["delta"][0, 0] = 20.0, 20.0, 4.0
    a["x"][0, 1], a["kind"][0, 1] = 80.0, 3
    a["check_collisions"][0] = 1
    st = ob._state(0)
    no.lib().net_substeps(C.byref(graph), C.byref(cfg), C.byref(st), 10 * K.FPS)
    assert not a["crashed"][0, 0]
    assert a["x"][0, 0] == pytest.approx(80.0 - 10.0, abs=1) and a["y"][0, 0] == pytest.approx(0)
    assert a["speed"][0, 0] == pytest.approx(0, abs=1) and a["heading"][0, 0] == pytest.approx(0)


def test_dynamics_collision_with_obstacle():
    """vehicle/test_dynamics.py:56-60: a vehicle at [20, 0] and an Obstacle at [23, 0] -> both crashed"""
    graph, cfg, ob = _single_lane_oracle(2)
    a = ob.a
    a["x"][0, 0], a["speed"][0, 0], a["target_speed"][0, 0], a["delta"][0, 0] = 20.0, 10.0, 10.0, 4.0
    a["x"][0, 1], a["kind"][0, 1] = 23.0, 3
    a["check_collisions"][0] = 1
    st = ob._state(0)
    no.lib().net_substeps(C.byref(graph), C.byref(cfg), C.byref(st), 1)
    assert a["crashed"][0, 0] and a["crashed"][0, 1]
