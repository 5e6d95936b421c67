"""Named parity cases shared by oracle/make_golden.py and tests/ (TEST INFRASTRUCTURE ONLY).

Each case is (config, steps, noise_kind, oracle kwargs).  The first five are the
BASELINE.json configurations; the rest cover the reference's other
views / masks / noises / reward modes and its edge cases (tasks "-1" and multi-digit,
keep_objects_num=False terminal-by-exhaustion, V=5 and V=9, transparent observation).
"""
from maexp_b200.config import baseline_config, make_config

_THREE = dict(name="three_rooms", SIZE_X=25, SIZE_Y=25)


def cases():
    c = {}
    c["cfg1_simple_mlp"] = (baseline_config(1), 60, None, {})
    c["cfg2_four_rooms_transition_flip_sharedmean"] = (baseline_config(2), 60, None, {})
    c["cfg3_three_rooms"] = (baseline_config(3), 60, None, {})
    c["cfg4_four_rectangle_merged_types"] = (baseline_config(4), 40, None, {})
    c["cfg5_four_rooms_n8"] = (baseline_config(5), 230, None, {})
    c["simple_views_v5"] = (make_config(agent_view_method="simple", object_view_method="simple", visible_range=5),
                            40, None, {})
    c["indtrans_t2_v9_tasks"] = (make_config(
        agent_view_method="individual_transition", past_step=2, object_view_method="object_types", type_objects=2,
        agent_tasks=["01", "10", "0", "1", "-1", "01", "0", "1"], visible_range=9,
        flip_noise_probabilities=[0.2, 0.1, 0.05, 0.02]), 40, None, {})
    c["relative_transparent"] = (make_config(observation_area_mask="relative", transparent_observation=True,
                                             num_agents=4), 40, None, {})
    c["sensing_dist"] = (make_config(observation_noise="sensing_dist", num_agents=3), 30, "normal", {})
    c["threshold_sensing_dist"] = (make_config(observation_noise="threshold_sensing_dist", num_agents=3,
                                               noise_threshold_ratio=0.7), 30, "normal", {})
    c["flat"] = (make_config(observation_noise="flat", num_agents=3), 30, "normal", {})
    c["merged_flip_three_rooms"] = (make_config(observation_noise="flip", observation_area_mask="merged",
                                                num_agents=3, map=_THREE), 40, None, {})
    c["no_keep_small"] = (make_config(keep_objects_num=False, num_objects=3, num_agents=6, max_episode_length=400,
                                      map=dict(name="simple", SIZE_X=10, SIZE_Y=10),
                                      init_xys=[[1, 1], [-1, 1], [1, -1], [-1, -1], [2, 2], [-2, -2]]),
                          300, None, {})
    c["shared_max_reference_compat"] = (make_config(environment="shared_max_reward", num_agents=5,
                                                    shuffle_init_xys=False), 60, None,
                                        {"shared_max_is_mean": True})
    # ---- sequential / junctional environments and the central-room maps
    xy8 = [[-2, 2], [2, 2], [-2, -2], [2, -2], [0, 2], [0, -2], [2, 0], [-2, 0]]  # configs/experiments/IJCNN2025/exp2
    c["sequential_central_room_t3"] = (make_config(
        environment="sequential", map=dict(name="central_room", SIZE_X=25, SIZE_Y=25), type_objects=3, num_objects=60,
        agent_tasks=["0", "0", "1", "1", "2", "2", "012", "01"], init_xys=xy8, agent_view_method="individual",
        object_view_method="object_types"), 160, None, {})
    c["junctional_central_room_t5_merged"] = (make_config(
        environment="junctional", map=dict(name="central_room", SIZE_X=25, SIZE_Y=25), type_objects=5, num_objects=60,
        agent_tasks=["0", "0", "12", "12", "13", "13", "4", "4"], init_xys=xy8, agent_view_method="individual",
        object_view_method="object_types", observation_area_mask="merged"), 160, None, {})
    c["junctional_chain_tasks"] = (make_config(
        environment="junctional", map=dict(name="central_room", SIZE_X=25, SIZE_Y=25), type_objects=5, num_objects=80,
        agent_tasks=["01", "01234", "0123", "012", "0124", "01", "43210", "0"], init_xys=xy8, max_episode_length=90), 200, None, {})
    c["sequential_no_keep"] = (make_config(
        environment="sequential", map=dict(name="central_room", SIZE_X=25, SIZE_Y=25), type_objects=2, num_objects=12,
        keep_objects_num=False, agent_tasks=["01"] * 8, init_xys=xy8, max_episode_length=400), 300, None, {})
    c["central_room_destination"] = (make_config(
        map=dict(name="central_room_destination", SIZE_X=25, SIZE_Y=25), num_agents=6, init_xys=xy8[:6], num_objects=60,
        max_episode_length=30), 100, None, {})
    c["central_room_large_destination_v9"] = (make_config(
        map=dict(name="central_room_large_destination", SIZE_X=49, SIZE_Y=25), num_agents=6, init_xys=xy8[:6], num_objects=90,
        visible_range=9, flip_noise_probabilities=[0.2, 0.1, 0.05, 0.02], max_episode_length=30), 70, None, {})
    c["central_room_large_sequential"] = (make_config(
        environment="sequential", map=dict(name="central_room_large", SIZE_X=49, SIZE_Y=25), type_objects=2, num_objects=90,
        agent_tasks=["0", "1", "01", "10", "0", "1", "01", "10"], init_xys=xy8), 80, None, {})
    c["two_central_rooms_t2_n12"] = (make_config(
        map=dict(name="two_central_rooms", SIZE_X=50, SIZE_Y=25), type_objects=2, num_agents=12, num_objects=60,
        agent_tasks=["0", "0", "1", "1", "01", "01", "10", "10", "0", "1", "01", "-1"],
        init_xys=xy8 + [[-3, 3], [3, 3], [-3, -3], [3, -3]], object_view_method="object_types"), 80, None, {})
    # dynamic_simple: two one-cell obstacles inside the object area (configs/map/dynamic_simple.yaml, dra3_exp.yaml);
    # the 10x10 variant crowds six agents and objects around them
    c["dynamic_simple_20"] = (make_config(map=dict(name="dynamic_simple", SIZE_X=20, SIZE_Y=20), num_agents=4,
                                          num_objects=40), 80, None, {})
    c["dynamic_simple_small_t2"] = (make_config(
        map=dict(name="dynamic_simple", SIZE_X=10, SIZE_Y=10), type_objects=2, num_objects=8, num_agents=6,
        agent_tasks=["0", "1", "01", "10", "0", "1"], object_view_method="object_types", max_episode_length=50,
        init_xys=[[1, 1], [-1, 1], [1, -1], [-1, -1], [2, 2], [-2, -2]]), 120, None, {})
    # default_environment.py:33,97-98 with more init_xys than agents: all 12 entries are shuffled, 5 are used
    c["shuffle_spare_init_xys"] = (make_config(num_agents=5, init_xys=xy8 + [[-3, 3], [3, 3], [-3, -3], [3, -3]],
                                               map=dict(name="central_room", SIZE_X=25, SIZE_Y=25), num_objects=40,
                                               max_episode_length=12), 80, None, {})
    return c
