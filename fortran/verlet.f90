!******************************************************************************************************
! Drop-in replacement for scripts/physics/verlet.f90: same module name, same procedure name and
! dummy list (verlet.f90:41-48).
!******************************************************************************************************
module verlet
  use define_precision,    only: dp_kind, int_kind
  use md_types,            only: sim_params, sim_state
  use ljmd_c_api
  implicit none

contains

  subroutine verlet_step(params, state, epot, ekin, d_epot, dd_epot)
    type(sim_params), intent(in)    :: params
    type(sim_state),  intent(inout) :: state
    real(kind=dp_kind), intent(out) :: epot, ekin, d_epot, dd_epot
    type(ljmd_params) :: cp

    if (params%n <= 0_int_kind) stop 'verlet_step(): params%n must be > 0.'
    if (.not. allocated(state%rx)) stop 'verlet_step(): state arrays are not allocated.'
    cp = to_c_params(params%n, params%num_cells, params%box_length, params%inv_box_length, &
                     params%volume, params%density, params%dt, params%dt_half, params%dt_square_half, &
                     params%rc, params%rc_square, params%length_scale, params%inv_length_scale)
    call ljmd_stop_on_error(ljmd_verlet_step(cp, state%rx, state%ry, state%rz, &
                                             state%vx, state%vy, state%vz, &
                                             state%ax, state%ay, state%az, epot, ekin, d_epot, dd_epot))
  end subroutine verlet_step

end module verlet
