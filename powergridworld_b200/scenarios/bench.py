"""BASELINE.json's benchmark configurations as ready-made product envs (bench.py, smoke(),
tools/): C1 = IEEE-13 coordinated buildings (configs[1]), C2 = component-only EV station + PV
+ storage (configs[2]), C3 = 123-bus-class feeder with 100 DER agents (configs[3]), HS = the
fork's Home-Steward house.  ``ns`` swaps the plugin classes (the CPU arm passes the oracle's)."""
import warnings

from powergridworld_b200.scenarios import catalog, catalog_hs

C1_LOAD_FACTOR = 1.2          # examples/marl/openai/train.py:165-188


def _ns(ns):
    if ns is None:
        from powergridworld_b200.scenarios.namespace import PRODUCT_NS
        return PRODUCT_NS
    return ns


def c1_env(ns=None, **kw):
    ns = _ns(ns)
    return ns.CoordinatedMultiBuildingControlEnv(
        **catalog.buildings_scenario(ns, ns.OpenDSSSolver, C1_LOAD_FACTOR), **kw)


def c2_env(ns=None, pf_cls=None, **kw):
    ns = _ns(ns)
    return ns.MultiAgentEnv(**catalog.ev_pv_storage_scenario(ns, pf_cls), **kw)


def c3_env(ns=None, **kw):
    ns = _ns(ns)
    with warnings.catch_warnings():          # agents on non-PQ loads are ignored, as in the reference
        warnings.simplefilter("ignore")
        return ns.MultiAgentEnv(**catalog.der123_scenario(ns, ns.OpenDSSSolver), **kw)


def hs_env(**kw):
    from powergridworld_b200.base_hs import house_agent_config
    from powergridworld_b200.scenarios.namespace import PRODUCT_HS_NS as HNS, PRODUCT_NS as NS
    cfg = catalog_hs.shipped(HNS)
    return NS.MultiAgentEnv(
        common_config={"start_time": cfg["start_time"], "end_time": "01-01-2031 00:00:00",
                       "control_timedelta": cfg["control_timedelta"]},
        pf_config=None, agents=[{"name": "house", "bus": None, "cls": HNS.HSMultiComponentEnv,
                                 "config": house_agent_config(cfg)}], **kw)


def make_env(workload: str, ns=None, **kw):
    return {"c1": c1_env, "c2": c2_env, "c3": c3_env}[workload](ns, **kw) if workload != "hs" \
        else hs_env(**kw)
