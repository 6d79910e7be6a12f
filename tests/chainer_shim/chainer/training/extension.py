class Extension:
    trigger = (1, "iteration")
    priority = 100
